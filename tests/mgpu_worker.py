"""torchrun worker for test_two_gpu_sharded_update_matches_single_gpu: the point-sharded NCCL path on 2 ranks must
give the single-GPU answer (N_eff identical; reduced system to FP64 rounding; final state within 1e-9)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ma-lio_b200"))
from malio_b200 import dist as mdist, plugin, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
case = synth.make_case("mgpu", 20000, 200000, 3, 3, varied_map_cov=True)
snap = plugin.build_static_snapshot(case.map_xyz, case.map_normal_y)
lo, hi = mdist.shard_bounds(case.pts.shape[0], rank, world)
m = plugin.MeasurementModel(3, device=local, params=case.params)
mdist.init_comm(m, rank, world, device=torch.device("cuda", local))
m.upload_map(snap)
m.upload_scan(case.pts[lo:hi], case.table, case.table_off, case.temporal_comp)
ok, HTH, HTh, st = m.h_share_model(case.x_prop, True)
x, P = case.x_prop.copy(), case.P_prop.copy()
m.rearm_scan()
rep = m.update_iterated_dyn_share_modified(x, P, 3)
vec = synth.state_to_vec(x, 3)
# every rank must hold the identical result
t = torch.from_numpy(np.concatenate([HTH.ravel(), vec])).cuda()
t0 = t.clone()
dist.broadcast(t0, 0)
assert torch.equal(t, t0), "ranks disagree"
if rank == 0:
    s = plugin.MeasurementModel(3, device=local, params=case.params)
    s.upload_map(snap)
    s.upload_scan(case.pts, case.table, case.table_off, case.temporal_comp)
    ok1, HTH1, HTh1, st1 = s.h_share_model(case.x_prop, True)
    assert st1.n_eff == st.n_eff, (st1.n_eff, st.n_eff)
    assert abs(st1.loc_weight - st.loc_weight) < 1e-12
    assert np.abs(HTH - HTH1).max() / np.abs(HTH1).max() < 1e-11
    x1, P1 = case.x_prop.copy(), case.P_prop.copy()
    s.rearm_scan()
    rep1 = s.update_iterated_dyn_share_modified(x1, P1, 3)
    assert rep1.passes == rep.passes and rep1.searches == rep.searches
    assert np.abs(synth.state_to_vec(x1, 3) - vec).max() < 1e-9
    print("MGPU_OK n_eff", st.n_eff, "passes", rep.passes)

# degenerate branch (esekfom.hpp:574-582, N_eff < n = 35): every rank must build H from the rows of ALL ranks
few = np.ascontiguousarray(case.pts[:: max(case.pts.shape[0] // 26, 1)][:26])
flo, fhi = mdist.shard_bounds(few.shape[0], rank, world)
m.upload_scan(few[flo:fhi], case.table, case.table_off, case.temporal_comp)
xd, Pd = case.x_prop.copy(), case.P_prop.copy()
repd = m.update_iterated_dyn_share_modified(xd, Pd, 3)
vd = synth.state_to_vec(xd, 3)
td = torch.from_numpy(np.concatenate([vd, Pd.ravel()])).cuda()
td0 = td.clone()
dist.broadcast(td0, 0)
assert torch.equal(td, td0), "ranks disagree in the degenerate branch"
if rank == 0:
    assert 0 < repd.n_eff_last < 35, repd.n_eff_last
    s.upload_scan(few, case.table, case.table_off, case.temporal_comp)
    x2, P2 = case.x_prop.copy(), case.P_prop.copy()
    rep2 = s.update_iterated_dyn_share_modified(x2, P2, 3)
    assert rep2.n_eff_last == repd.n_eff_last and rep2.passes == repd.passes
    assert np.abs(synth.state_to_vec(x2, 3) - vd).max() < 1e-9
    assert np.abs(P2 - Pd).max() / np.abs(P2).max() < 1e-9
    print("MGPU_DEGENERATE_OK n_eff", repd.n_eff_last)
dist.barrier()
m.close()
dist.destroy_process_group()
