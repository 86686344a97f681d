python - <<'PY'
import sys, time, torch
sys.path.insert(0, '.')
import bench
from dcrmontecarlo_b200 import scenarios as sc, _native as nat
s = sc.cfg4(); solver = s.make_solver()
pts_h = bench.tile_points(s.points, 16384); pts_d = pts_h.cuda()
for i in range(6): solver.solve_raw(pts_d, 64, s.max_steps, s.eps, seed=100+i, device_outputs=True)
torch.cuda.synchronize()
pin = pts_h.pin_memory()
for i in range(6): solver.solve_raw(pin, 64, s.max_steps, s.eps, seed=200+i)
sp_d = s.points.cuda()
ts = []
for i in range(60):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = solver.solve_raw(sp_d, 25, s.max_steps, s.eps, seed=300+i, device_outputs=True)
    torch.cuda.synchronize(); ts.append((time.perf_counter()-t0)*1e6)
print("per-solve us:", [int(t) for t in ts])
print(nat.jit_stats())
PY
