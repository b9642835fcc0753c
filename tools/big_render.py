"""A render twice the size of BASELINE's configs[4] through the plugin API: 512 x 180 s stereo tracks, three mix levels
(32 groups of 16 -> 2 mixes of 16 -> master), sources resident in HBM (32.5 GB).  Runs under both memory policies
(keep / release, infra::Runner::release_products), checks the end tracks' chains against the oracle and the master bus
against the ordered sum of its inputs, and prints one JSON line with times and the allocator's high-water marks.
usage (GPU box): python tools/big_render.py   (TRACKS=512 SECONDS=180 by default)"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nodey-audio-editor_b200", "bindings"))

import numpy as np
import torch

import engine
import nodey
import pipeline

TRACKS = int(os.environ.get("TRACKS", "512"))
SECONDS = int(os.environ.get("SECONDS", "180"))
RATE = 44100


def project(n, gains):
    assert n % 256 == 0
    p = engine.Project()
    src = p.add("audio_input", {"file_path": [""] * n})
    gain_nodes = []
    for t in range(n):
        a = p.add("audio_amix", engine.amix_info([1.0]))
        pm = p.add("pitch_modifier", {"pitch": 3.0})
        vm = p.add("velocity_modifier", {"velocity": 1.25, "keep_pitch": True})
        g = p.add("audio_volume_adjust", {"volume": gains[t]})
        p.link(src, f"output_{t}", a, "input_1"); p.link(a, "output", pm, "input")
        p.link(pm, "output", vm, "input"); p.link(vm, "output", g, "input")
        gain_nodes.append(g)

    def mix_level(nodes, vol):
        out = []
        for k in range(0, len(nodes), 16):
            part = nodes[k:k + 16]
            m = p.add("audio_amix", engine.amix_info([vol] * len(part)))
            for j, node in enumerate(part):
                p.link(node, "output", m, f"input_{j + 1}")
            out.append(m)
        return out
    groups = mix_level(gain_nodes, 1.0 / 16)
    supers = mix_level(groups, 1.0 / 16)
    master = p.add("audio_amix", engine.amix_info([1.0 / len(supers)] * len(supers)))
    for j, node in enumerate(supers):
        p.link(node, "output", master, f"input_{j + 1}")
    out = p.add("audio_output")
    p.link(master, "output", out, "input")
    return p, {"gains": gain_nodes, "groups": groups, "supers": supers, "master": master}


def main():
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    L = nodey.lib()
    engine.lib()
    n_in = RATE * SECONDS
    gains = [pipeline.track_gain(t) for t in range(TRACKS)]
    p, ids = project(TRACKS, gains)
    eng = engine.Engine(p.json())
    x = torch.empty((TRACKS, n_in, 2), dtype=torch.float32, device=dev)
    for t in range(TRACKS):
        nodey.check(L.nodey_synth(nodey._dp(x[t]), None, n_in, 2, RATE, t, 0, None))
    torch.cuda.synchronize()
    for t in range(TRACKS):
        eng.bind_source(t, x[t], nodey.FMT_FLT, RATE)
    res = {"tracks": TRACKS, "seconds": SECONDS, "sources_gb": round(x.numel() * 4 / 1e9, 2)}
    buses = {}
    for policy in ("release", "keep"):
        engine.set_release_products(policy == "release")
        nodey.check(L.nodey_trim_memory())
        nodey.memory_stats(reset_peak=True)
        t0 = time.perf_counter(); eng.run(); torch.cuda.synchronize(); cold = (time.perf_counter() - t0) * 1e3
        _, peak = nodey.memory_reserved()
        t0 = time.perf_counter(); eng.run(); torch.cuda.synchronize(); warm = (time.perf_counter() - t0) * 1e3
        buses[policy] = eng.output().numpy().copy()
        res[policy] = {"cold_run_ms": round(cold, 1), "warm_run_ms": round(warm, 1), "peak_gb": round(peak / 1e9, 2),
                       "audio_s_per_s": round(TRACKS * SECONDS / (warm / 1e3), 1)}
    engine.set_release_products(False)
    # parity (products are still there: the last runs kept them)
    from oracle import graph_oracle as G
    from oracle import oracle as O
    O.build()
    ok = bool(np.array_equal(buses["keep"].view(np.uint32), buses["release"].view(np.uint32)))
    res["release_equals_keep"] = ok
    for t in (0, TRACKS - 1):
        ref = G.track_chain(x[t].cpu().numpy(), G.track_gain(t))
        got = eng.product(ids["gains"][t], "output").numpy()
        same = got.shape == ref.shape and bool(np.array_equal(got.view(np.uint32), ref.view(np.uint32)))
        res[f"track_{t}_bit_exact"] = same
        ok = ok and same

    def mixed(nodes, vol, frames):
        acc = np.zeros((2, frames), np.float32)
        v = np.float32(vol)
        for node in nodes:
            gm = eng.product(node, "output").numpy()
            acc[:, :gm.shape[1]] = acc[:, :gm.shape[1]] + gm * v
        return acc
    bus = buses["keep"]
    same = bool(np.array_equal(bus.view(np.uint32), mixed(ids["supers"], 1.0 / len(ids["supers"]), bus.shape[1]).view(np.uint32)))
    res["master_bus_bit_exact"] = same
    ok = ok and same
    s0 = eng.product(ids["supers"][0], "output").numpy()
    same = bool(np.array_equal(s0.view(np.uint32), mixed(ids["groups"][:16], 1.0 / 16, s0.shape[1]).view(np.uint32)))
    res["second_level_mix_bit_exact"] = same
    res["ok"] = bool(ok and same)
    print(json.dumps(res))
    eng.close()
    if not res["ok"]:
        raise SystemExit("big render: parity failed")


if __name__ == "__main__":
    main()
