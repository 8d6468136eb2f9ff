"""Locality statistics of the lane search on the C4 workload (host emulation; design aid).
usage: python tools/lane_stats.py [scene] [queries]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from conftest import hostemu, navmesh_image
from workloads.scenes import NavMeshGeom, pointnav_pairs
name = sys.argv[1] if len(sys.argv) > 1 else "c4_building"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
emu = hostemu()
img = navmesh_image(name)
h = C.c_void_p(emu.emu_create(img, C.c_long(len(img))))
st, en = pointnav_pairs(NavMeshGeom(img), n, 17)
st = np.ascontiguousarray(st, np.float32); en = np.ascontiguousarray(en, np.float32)
out = (C.c_long * 32)()
f32p = C.POINTER(C.c_float)
emu.emu_lane_stats(h, st.ctypes.data_as(f32p), en.ctypes.data_as(f32p), C.c_long(n), out)
o = list(out)
print(f"searches {o[0]} expansions/search {o[1]/o[0]:.1f} nodes/search {o[2]/o[0]:.1f}")
print(f"u16 table sectors/search {o[3]/o[0]:.1f}  bitmap sectors/search {o[4]/o[0]:.2f}  bitmap span {o[5]/o[0]*16:.0f} B/search")
tot = sum(o[6:18]); acc = 0
for b in range(12):
    acc += o[6 + b]
    print(f"pop age <= {1 << b if b < 11 else 'inf'}: {acc/tot*100:.1f} %")
print(f"distinct key blocks per search: 8-key mean {o[18]/o[0]:.1f} max {o[21]}, 16-key mean {o[19]/o[0]:.1f} max {o[22]}, 32-key mean {o[20]/o[0]:.1f} max {o[23]}; searches over 256 16-key blocks: {o[24]}")
