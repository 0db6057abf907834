"""The *ELEMENT emitter of one slab at the offsets rank R of an N-rank job sees (ids beyond 2^32 take the 64-bit
variant of k_emit_elements2): cfe_fill_elements / cfe_element_lengths / cfe_emit_elements per entry point.
python profiles/tools/bench_emit_wide.py [rank] [Z,Y,X]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ciclope_b200 import _lib, synth, vol2ugrid
from ciclope_b200.voxelFE import _ELEM_HDR_SLOT

rank = int(sys.argv[1]) if len(sys.argv) > 1 else 7
shape = tuple(int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "256,2048,2048").split(","))
vol = synth.trabecular(shape, 0, 0.30, 'cuda')
m = vol2ugrid(vol, [0.0195] * 3)
Z, Y, X = shape
W = (X + 31) // 32
z0, eid0 = rank * Z, rank * m.n_el
hdr = bytearray(256 * _ELEM_HDR_SLOT)
h = b'*ELEMENT, TYPE=C3D8, ELSET=SET1\n'
hdr[_ELEM_HDR_SLOT:_ELEM_HDR_SLOT + len(h)] = h
hlen = np.zeros(256, np.uint8); hlen[1] = len(h) if rank == 0 else 0
first = np.zeros(256, np.int64); first[2:] = m.n_el
blob = torch.from_numpy(np.frombuffer(bytes(hdr) + hlen.tobytes() + first.tobytes(), dtype=np.uint8).copy()).cuda()
base = blob.data_ptr()
o_hlen, o_first = 256 * _ELEM_HDR_SLOT, 256 * _ELEM_HDR_SLOT + 256
out = torch.empty(m.n_el * 112 + 64, dtype=torch.uint8, device='cuda')
totals = torch.zeros(2, dtype=torch.int64, device='cuda')
s = _lib.stream()
el_args = (m.elem_vox.data_ptr(), None, m.n_el, vol.data_ptr(), base + o_first, base, base + o_hlen, 1, Z, Y, X, z0, eid0)


def step():
    _lib.call("cfe_fill_elements_gv", m.bits.data_ptr(), m.epref.data_ptr(), Z * Y * W, Y, X, z0, eid0, m.n_el,
              m.elem_vox.data_ptr(), m.ws_el.data_ptr(), vol.data_ptr(), None, s)
    _lib.call("cfe_element_lengths", *el_args, 1, totals.data_ptr(), m.ws_el.data_ptr(), s)
    _lib.call("cfe_emit_elements", *el_args, out.data_ptr(), None, m.ws_el.data_ptr(), s)


for _ in range(3):
    step()
torch.cuda.synchronize()
_lib.PROFILE = []
for _ in range(5):
    step()
torch.cuda.synchronize()
ms = {}
for name, a, b in _lib.PROFILE:
    ms[name] = ms.get(name, 0.0) + a.elapsed_time(b) / 5
nbytes = int(totals[0].item())
print(os.environ.get("CICLOPE_B200_LIB", "default"), "rank", rank, "bytes", nbytes, {k: round(v, 3) for k, v in ms.items()},
      "emit GB/s %.0f" % (nbytes / ms["cfe_emit_elements"] / 1e6))
