"""CPU: the C++ host layer's graph model and project JSON (reference: src/infra/graph.cpp) through the
C facade -- registration, pins, (de)serialisation round trip, validation errors, graph levels.
No kernel is launched (engines are created and inspected, never run)."""
import json

import pytest


@pytest.fixture(scope="module")
def eng():
    import engine
    engine.lib()
    return engine


def simple_project(eng, n_inputs=2):
    p = eng.Project()
    src = p.add("audio_input", {"file_path": ["a.wav", "b.wav"]})
    g0 = p.add("audio_volume_adjust")
    g1 = p.add("audio_volume_adjust")
    mix = p.add("audio_amix", eng.amix_info([0.5] * n_inputs))
    out = p.add("audio_output")
    p.link(src, "output_0", g0, "input")
    p.link(src, "output_1", g1, "input")
    p.link(g0, "output", mix, "input_1")
    p.link(g1, "output", mix, "input_2")
    p.link(mix, "output", out, "input")
    return p


def test_project_round_trip(eng):
    p = simple_project(eng)
    e = eng.Engine(p.json())
    text = e.serialize()
    again = json.loads(text)
    assert sorted(again["nodes"]) == ["0", "1", "2", "3", "4"]
    assert again["nodes"]["0"] == {"identifier": "audio_input", "info": {"file_path": ["a.wav", "b.wav"]}, "position": {"x": 0.0, "y": 0.0}}
    assert again["nodes"]["1"]["info"] is None                   # gain is not serialised (App. C1)
    assert again["nodes"]["3"]["info"] == {"input_num": 2, "volumes0": 0.5, "locks0": False, "volumes1": 0.5, "locks1": False}
    assert {json.dumps(l, sort_keys=True) for l in again["links"]} == {json.dumps(l, sort_keys=True) for l in p.json()["links"]}
    e2 = eng.Engine(text)
    assert json.loads(e2.serialize()) == again
    assert text.startswith("{\n  \"links\"")                      # 2-space indent, keys in order


def test_nodes_levels_and_identifiers(eng):
    e = eng.Engine(simple_project(eng).json())
    e.check()
    assert e.nodes() == [(0, "audio_input", 0), (1, "audio_volume_adjust", 1), (2, "audio_volume_adjust", 1),
                         (3, "audio_amix", 2), (4, "audio_output", 3)]


@pytest.mark.parametrize("ident", ["audio_input", "audio_output", "audio_volume_adjust", "velocity_modifier", "pitch_modifier",
                                   "audio_amix", "audio_bimix", "audio_bimix_v2", "audio_channel_split", "audio_spectrum"])
def test_every_identifier_is_registered(eng, ident):
    p = eng.Project()
    info = {"audio_input": {"file_path": [""]}, "audio_amix": eng.amix_info([1.0]), "audio_bimix": {"bias": 0.0}}.get(ident)
    p.add(ident, info)
    e = eng.Engine(p.json())
    assert e.nodes()[0][1] == ident


def test_invalid_files(eng):
    def err(doc):
        with pytest.raises(eng.EngineError) as x:
            eng.Engine(doc if isinstance(doc, str) else json.dumps(doc))
        return x.value
    assert err("{ not json").code == eng.E_FILE
    assert err([1, 2]).code == eng.E_FILE
    assert err({"nodes": [], "links": []}).code == eng.E_FILE
    assert err({"nodes": {}, "links": {}}).code == eng.E_FILE
    assert "Unknown processor identifier" in err({"nodes": {"0": {"identifier": "nope", "info": None}}, "links": []}).message
    assert "Invalid node ID" in err({"nodes": {"1x": {"identifier": "audio_output", "info": None}}, "links": []}).message
    two = {"nodes": {"0": {"identifier": "audio_output", "info": None}, "1": {"identifier": "audio_output", "info": None}}, "links": []}
    assert "Duplicating singleton" in err(two).message
    p = simple_project(eng).json()
    p["links"].append({"from": {"node": 9, "pin": "output"}, "to": {"node": 4, "pin": "input"}})
    assert "non-existent node" in err(p).message
    p = simple_project(eng).json()
    p["links"][0]["to"]["pin"] = "nope"
    assert "non-existent pin" in err(p).message
    assert "Wrong field: input_num" in err({"nodes": {"0": {"identifier": "audio_amix", "info": {}}}, "links": []}).message
    assert "Wrong field: bias" in err({"nodes": {"0": {"identifier": "audio_bimix", "info": {"bias": "x"}}}, "links": []}).message
    assert "Wrong field: file_path" in err({"nodes": {"0": {"identifier": "audio_input", "info": {}}}, "links": []}).message


def test_multiple_input_and_type_mismatch(eng):
    p = simple_project(eng).json()
    p["links"].append({"from": {"node": 2, "pin": "output"}, "to": {"node": 3, "pin": "input_1"}})   # second link into input_1
    # like the reference, add_link() only refuses a link when the pin ALREADY has two (graph.hpp:173-183
    # counts the existing links); the doubled input is caught by check_graph()
    e = eng.Engine(p)
    with pytest.raises(eng.EngineError) as x:
        e.check()
    assert x.value.code == eng.E_GRAPH and "Multiple Inputs" in x.value.message
    q = eng.Project()
    s = q.add("audio_spectrum", {"fft_size": 4096, "hop": 1024, "window": "hann"})
    o = q.add("audio_output")
    q.link(s, "output", o, "input")        # spectrum product into an audio pin
    with pytest.raises(eng.EngineError) as x:
        eng.Engine(q.json())
    assert x.value.code == eng.E_GRAPH and "Mismatch Pin" in x.value.message


def test_loop_detection(eng):
    p = eng.Project()
    a = p.add("audio_volume_adjust")
    b = p.add("audio_volume_adjust")
    p.link(a, "output", b, "input")
    p.link(b, "output", a, "input")
    e = eng.Engine(p.json())
    with pytest.raises(eng.EngineError) as x:
        e.check()
    assert "Loop Detected" in x.value.message
    # a loop hanging off a valid source
    q = eng.Project()
    s = q.add("audio_input", {"file_path": [""]})
    m = q.add("audio_amix", eng.amix_info([1.0, 1.0]))
    g = q.add("audio_volume_adjust")
    q.link(s, "output_0", m, "input_1")
    q.link(m, "output", g, "input")
    q.link(g, "output", m, "input_2")
    with pytest.raises(eng.EngineError):
        eng.Engine(q.json()).check()


def test_amix_pins_follow_input_num(eng):
    p = eng.Project()
    m = p.add("audio_amix", eng.amix_info([1.0] * 16))
    g = p.add("audio_volume_adjust")
    p.link(g, "output", m, "input_16")
    eng.Engine(p.json())
    p2 = eng.Project()
    m = p2.add("audio_amix", eng.amix_info([1.0] * 3))
    g = p2.add("audio_volume_adjust")
    p2.link(g, "output", m, "input_4")
    with pytest.raises(eng.EngineError):
        eng.Engine(p2.json())


def test_config5_project_shape(eng):
    p, ids = eng.config5_project(32, [1.0] * 32)
    e = eng.Engine(p.json())
    e.check()
    nodes = e.nodes()
    assert len(nodes) == 1 + 32 * 4 + 2 + 1 + 1 + 1
    levels = {nid: lvl for nid, _, lvl in nodes}
    assert levels[ids["input"]] == 0 and levels[ids["master"]] == 6 and levels[ids["output"]] == 7


def test_compat_switches_are_written_only_when_set(eng):
    """the optional JSON keys of the quirk switches: absent from serialised projects unless
    set, so files written by the reference round-trip unchanged; set, they survive a round trip"""
    def build(flags):
        p = eng.Project()
        src = p.add("audio_input", {"file_path": ["", ""]})
        mix = p.add("audio_amix", dict(eng.amix_info([1.0, 1.0]), **({"start_time_stamps": True} if flags else {})))
        bi = p.add("audio_bimix", dict({"bias": 0.0}, **({"start_time_stamps": True} if flags else {})))
        pm = p.add("pitch_modifier", dict({"pitch": 1.0}, **({"reference_schedule": True} if flags else {})))
        out = p.add("audio_output")
        p.link(src, "output_0", mix, "input_1"); p.link(src, "output_1", mix, "input_2")
        p.link(mix, "output", pm, "input"); p.link(pm, "output", out, "input")
        p.link(src, "output_0", bi, "input_l"); p.link(src, "output_1", bi, "input_r")
        return p, mix, bi, pm
    for flags in (False, True):
        p, mix, bi, pm = build(flags)
        text = json.loads(eng.Engine(p.json()).serialize())
        again = json.loads(eng.Engine(json.dumps(text)).serialize())
        assert again == text
        for node, key in ((mix, "start_time_stamps"), (bi, "start_time_stamps"), (pm, "reference_schedule")):
            info = text["nodes"][str(node)]["info"]
            assert (info.get(key) is True) if flags else (key not in info)


def test_example_frame_processor_registers_and_round_trips():
    """frame-streaming compatibility mode: the frame-interface example node is outside the reference's set until the
    host registers it; afterwards it loads from project JSON like any other processor"""
    import engine
    engine.register_examples()
    engine.register_examples()          # idempotent
    p = engine.Project()
    src = p.add("audio_input", {"file_path": [""]})
    fg = p.add("frame_gain_example", {"volume": 0.25})
    out = p.add("audio_output")
    p.link(src, "output_0", fg, "input"); p.link(fg, "output", out, "input")
    e = engine.Engine(p.json())
    e.check()
    again = json.loads(e.serialize())
    assert again["nodes"][str(fg)]["identifier"] == "frame_gain_example"
    assert again["nodes"][str(fg)]["info"] == {"volume": 0.25}


def test_graph_editing_api_cpp(tmp_path):
    """the C++ editing API the editor drives (add_node / remove_node / update_node_pin / add_link / remove_link, id
    reuse, singleton map, links re-attached by pin name): tests/cpp/graph_edit_test.cpp against libnodey_host.so"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "nodey-audio-editor_b200")
    exe = str(tmp_path / "graph_edit_test")
    subprocess.run(["g++", "-std=c++20", "-O1", "-I" + os.path.join(pkg, "host", "shim"), "-I" + os.path.join(pkg, "host", "include"),
                    "-I" + os.path.join(root, "include"), "-o", exe, os.path.join(root, "tests", "cpp", "graph_edit_test.cpp"),
                    "-L" + pkg, "-lnodey_host", "-lnodey_cuda", "-Wl,-rpath," + pkg, "-lpthread"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    assert "graph_edit_test ok" in r.stdout


def test_schedule_and_frame_clock_cpp(tmp_path):
    """infra::Runner::Schedule (environment overrides read once per run, explicit settings win) and processor::Frame_clock
    (per-frame stamps by the producing node's rule): tests/cpp/schedule_clock_test.cpp against libnodey_host.so"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "nodey-audio-editor_b200")
    exe = str(tmp_path / "schedule_clock_test")
    subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-I" + os.path.join(pkg, "host", "shim"), "-I" + os.path.join(pkg, "host", "include"),
                    "-I" + os.path.join(root, "include"), "-o", exe, os.path.join(root, "tests", "cpp", "schedule_clock_test.cpp"),
                    "-L" + pkg, "-lnodey_host", "-lnodey_cuda", "-Wl,-rpath," + pkg, "-lpthread"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    assert "schedule_clock_test ok" in r.stdout


def test_node_faults_read_like_the_reference_cpp(tmp_path):
    """unlinked input pins and streams a node cannot take: Processor::Runtime_error with the reference's own message,
    explanation and detail, word for word (tests/cpp/node_errors_test.cpp against libnodey_host.so; no device needed)"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "nodey-audio-editor_b200")
    exe = str(tmp_path / "node_errors_test")
    subprocess.run(["g++", "-std=c++20", "-O1", "-I" + os.path.join(pkg, "host", "shim"), "-I" + os.path.join(pkg, "host", "include"),
                    "-I" + os.path.join(root, "include"), "-o", exe, os.path.join(root, "tests", "cpp", "node_errors_test.cpp"),
                    "-L" + pkg, "-lnodey_host", "-lnodey_cuda", "-Wl,-rpath," + pkg, "-lpthread"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    assert "node_errors_test ok" in r.stdout


def test_render_without_a_cuda_device_fails_loudly(eng):
    """no CPU fallback: on a box without a CUDA device the Runner marks every node as failed with a Runtime_error that
    says so (the product path never routes through the oracle)"""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a CUDA device")
    p = eng.Project()
    src = p.add("audio_input", {"file_path": [""]})
    g = p.add("audio_volume_adjust", {"volume": 0.5})
    out = p.add("audio_output")
    p.link(src, "output_0", g, "input"); p.link(g, "output", out, "input")
    e = eng.Engine(p.json())
    e.bind_source(0, np.zeros((1000, 2), np.float32), 3, 48000)
    with pytest.raises(eng.EngineError) as x:
        e.run()
    assert "No CUDA device" in x.value.message and "no CPU fallback" in x.value.message
    assert e.diagnostics().splitlines()[0] == "0 Running | 0 Finished | 3 Errors"


def test_malformed_projects_are_rejected_not_crashed(eng):
    """seeded fuzz of the project loader (json shim + Graph::deserialize + the nodes' deserialize): text-level damage
    and structural damage (dangling link ends, unknown identifiers, wrong-typed info, huge ids) end in EngineError or
    in a graph that still serialises -- never in a crash (a crash would take the test process down)"""
    import random
    rnd = random.Random(20261018)
    p, _ = eng.config5_project(16, [1.0] * 16)
    base = json.dumps(p.json())

    def damage_text(s):
        s = list(s)
        for _ in range(rnd.randint(1, 6)):
            k = rnd.randrange(len(s))
            op = rnd.random()
            if op < 0.3:
                del s[k]
            elif op < 0.6:
                s.insert(k, rnd.choice('{}[]",:0123456789.-eE\\ntrufalse \x00\xff'))
            elif op < 0.8:
                s[k] = rnd.choice('{}[]",:0123456789.-eE\\ \t\n')
            else:
                j = rnd.randrange(len(s))
                s[k:k] = s[j:j + rnd.randint(1, 40)]
        return "".join(s)

    def damage_structure(d):
        d = json.loads(json.dumps(d))
        nodes, links = d["nodes"], d["links"]
        r = rnd.random()
        if r < 0.2:
            rnd.choice(links)["from"]["node"] = rnd.choice([-1, 999999, "x", None, 1.5])
        elif r < 0.4:
            rnd.choice(links)["to"]["pin"] = rnd.choice(["", "input_99", 5, None])
        elif r < 0.6:
            nodes[rnd.choice(list(nodes))]["identifier"] = rnd.choice(["", "nope", 3, None, "audio_input"])
        elif r < 0.8:
            nodes[rnd.choice(list(nodes))]["info"] = rnd.choice([[], 5, "s", {"input_num": -5}, {"input_num": 1e30}, {"pitch": "x"},
                                                                  {"velocity": None}, {"file_path": 3}])
        else:
            nodes[str(rnd.choice([-1, 2 ** 31, 2 ** 40]))] = {"identifier": "audio_volume_adjust", "info": None, "position": {"x": "a"}}
        return json.dumps(d)

    loaded = rejected = 0
    for i in range(300):
        text = damage_text(base) if i % 2 else damage_structure(p.json())
        try:
            e = eng.Engine(text)
        except eng.EngineError:
            rejected += 1
            continue
        try:
            e.check()
        except eng.EngineError:
            pass
        json.loads(e.serialize())
        e.close()
        loaded += 1
    assert rejected > 150 and loaded + rejected == 300
