"""Developer probe: the same 200M-key build through rmi_train and through the one-call range-partitioned path with a
single rank (rmi_shard_train, world 1): per-phase device times side by side."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rmi_b200
from rmi_b200 import sharded

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000_000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
dev = torch.device("cuda:0")
torch.manual_seed(42)
k = torch.randint(0, 2**63 - 1, (n,), dtype=torch.int64, device=dev)
k, _ = torch.sort(k)
torch.cuda.synchronize()
ds = rmi_b200.RMITrainingData.from_device(k.data_ptr(), n, rmi_b200.KEY_U64, 0, keep_alive=k)
sd = sharded.ShardedTrainingData(k, n, rmi_b200.KEY_U64, halo_capacity=16)
out = {}
for name, fn in (("rmi_train", lambda: rmi_b200.train(ds, "linear,linear", N, 0, counts=False)),
                 ("rmi_train stats_only", lambda: rmi_b200.train(ds, "linear,linear", N, rmi_b200.FLAG_STATS_ONLY, counts=False)),
                 ("rmi_shard_train world 1", lambda: sharded.train_sharded(sd, "linear,linear", N, 0, counts=False, native=True)),
                 ("host-driven phases world 1", lambda: sharded.train_sharded(sd, "linear,linear", N, 0, counts=False, native=False))):
    walls, phases = [], None
    for it in range(8):
        t0 = time.perf_counter()
        r = fn()
        walls.append((time.perf_counter() - t0) * 1e3)
        p = [x / 1e6 for x in r.phase_device_ns]
        phases = p if phases is None or p[2] < phases[2] else phases
    out[name] = {"wall_ms_min": min(walls), "phases_ms_at_min_leaf": phases, "device_ms": r.device_time_ns / 1e6}
print(json.dumps(out))
