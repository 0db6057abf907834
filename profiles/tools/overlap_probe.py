"""Does an issue-bound kernel hide behind the HBM-bound *NODE emitter?  Times cfe_emit_nodes alone, the element-list
fill (+ line-length totals) alone, and both together with the fill on a high-priority side stream.
python profiles/tools/overlap_probe.py [Z,Y,X]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from ciclope_b200 import _lib, synth, vol2ugrid

shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256,2048,2048").split(","))
vol = synth.trabecular(shape, 0, 0.30, 'cuda')
m = vol2ugrid(vol, [0.0195] * 3)
Z, Y, X = m.Z, m.Y, m.X
W = (X + 31) // 32
rows = Z * Y
out = torch.empty(53 * m.n_nd + 64, dtype=torch.uint8, device='cuda')
out = out[(-out.data_ptr()) % 16:]
p_xt = m.ftab.data_ptr(); p_yt = p_xt + 16 * (X + 1); p_zt = p_yt + 16 * (Y + 1)
node_list = m.node_list
main = torch.cuda.current_stream()
lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
side_hi = torch.cuda.Stream(priority=-1)
side_eq = torch.cuda.Stream(priority=0)


def nodes(s):
    _lib.call("cfe_emit_nodes", node_list.data_ptr(), m.n_nd, p_xt, p_yt, p_zt, Y, X, 0, out.data_ptr(), s.cuda_stream)


def fill(s):
    _lib.call("cfe_fill_elements_gv", m.bits.data_ptr(), m.epref.data_ptr(), rows * W, Y, X, 0, 0, m.n_el,
              m.elem_vox.data_ptr(), m.ws_el.data_ptr(), vol.data_ptr(), None, s.cuda_stream)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(n):
        fn()
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def both(side, fill_first):
    def run():
        side.wait_stream(main)
        if fill_first:
            with torch.cuda.stream(side):
                fill(side)
            nodes(main)
        else:
            nodes(main)
            with torch.cuda.stream(side):
                fill(side)
        main.wait_stream(side)
    return run


print("nodes alone      %.3f ms" % timed(lambda: nodes(main)))
print("fill alone       %.3f ms" % timed(lambda: fill(main)))
print("serial           %.3f ms" % timed(lambda: (nodes(main), fill(main))))
print("side hi, fill 1st %.3f ms" % timed(both(side_hi, True)))
print("side hi, nodes 1st %.3f ms" % timed(both(side_hi, False)))
print("side eq, fill 1st %.3f ms" % timed(both(side_eq, True)))
print("side eq, nodes 1st %.3f ms" % timed(both(side_eq, False)))
