#!/bin/bash
set -x
./tools/micro/f32x2_check
python - <<'PY'
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "nodey-audio-editor_b200/bindings"); sys.path.insert(0, "tests")
import numpy as np, torch
import nodey as nd
from oracle import oracle as O
O.build()
x = O.synth_f32(48000, 2, 48000, 5)
st = nd.SoundTouch(48000, 2, 1.0, O.pitch_node_factor(3.0))
y = st.run(torch.from_numpy(x).cuda()).cpu().numpy()
ref, _, _ = O.soundtouch(x, 48000, 1.0, O.pitch_node_factor(3.0), 1152)
d = y.view(np.uint32) != ref.view(np.uint32)
print("mismatching samples", int(d.sum()), "of", d.size, "; of those both zero:", int((d & (y == 0) & (ref == 0)).sum()), "max abs diff", float(np.abs(y - ref).max()))
idx = np.argwhere(d)[:5]
for i, c in idx: print(i, c, y[i, c], ref[i, c], hex(y.view(np.uint32)[i, c]), hex(ref.view(np.uint32)[i, c]))
PY
