import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from ciclope_b200 import sharded as sh
from tests import cases
Zl, Y, X = 6, 48, 64
RN, SN = X + 1, (X + 1) * (Y + 1)
z0 = (1 << 32) // SN + 700_000
Zg = z0 + Zl
eid_base = 5_000_000_000 - 300
vs = [0.01, 0.02, 0.00003]
bw = cases.rand_binary((Zl, Y, X), 0.35, 77)
runner = sh.LocalRunner(1); ctx = runner.ctxs()[0]
slab = torch.from_numpy(bw).cuda().view(torch.uint8)
mesh = runner.run([sh.vol2ugrid_program(ctx, slab, True, Zg, z0, vs, eid_base=eid_base)])[0]
deck = runner.run([sh.deck_program(ctx, mesh, cases.TEMPLATE)])[0]
for off, buf, lo, hi in deck.writes:
    t = bytes(buf[lo:hi].cpu().numpy()).decode(errors='replace')
    print("piece", off, lo, hi, len(t)); print(t[:400]); print('...'); print(t[-300:])
print("n_el", mesh.n_el, "n_nd", mesh.n_nd, "eid_offset", mesh.eid_offset)
