"""Developer probe: plain rmi_train vs the phase API (world 1) on the same resident keys."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rmi_b200
from rmi_b200 import sharded
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000_000
dev = torch.device("cuda:0")
torch.manual_seed(1)
k = torch.randint(0, (2**63 - 1) // 2, (n,), dtype=torch.int64, device=dev)
k, _ = torch.sort(k)
buf = torch.empty(n + (1 << 20), dtype=torch.int64, device=dev)
buf[:n].copy_(k)
ds = rmi_b200.RMITrainingData.from_device(buf.data_ptr(), n, rmi_b200.KEY_U64, 0, keep_alive=buf)
sd = sharded.ShardedTrainingData(buf, n, rmi_b200.KEY_U64, halo_capacity=1 << 20)
for spec, N in [("linear,linear", 1 << 20), ("linear,linear", 1 << 19), ("radix,linear", 1 << 19)]:
    for it in range(3):
        a = rmi_b200.train(ds, spec, N)
        b = sharded.train_sharded(sd, spec, N)
    print(json.dumps({"spec": spec, "N": N, "plain_phases_ms": [p / 1e6 for p in a.phase_device_ns],
                      "shard_phases_ms": [p / 1e6 for p in b.phase_device_ns],
                      "max_leaf_keys": int(a.l1_counts.max()), "same_errors": bool((a.last_layer_max_l1s == b.last_layer_max_l1s).all())}))
