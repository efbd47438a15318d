"""Range-partitioned (multi-GPU) two-layer RMI build: one process per GPU, torch.distributed
for the three small collectives, the C ABI's rmi_shard_* phases for all arithmetic.

The result is the same TrainedRMI on every rank and equals a single-GPU build of the
concatenated key array (reference semantics: rmi_lib::train on the whole data set).

    data = ShardedTrainingData(local_sorted_keys_tensor, n_local)     # rank r holds the r-th slab
    rmi  = train_sharded(data, "linear,linear", 1 << 20)

Data path per build:
    top model        0, 1 or 2 tiny all-reduces (TOP_ROUNDS): SUM of 8 doubles (linear /
                     robust_linear sums; normal / lognormal mean, then variance; cubic L1
                     comparison), MIN of 4 x i64 (cubic: the two interior points of the spline)
    all-reduce MIN   (N+1) x u64        leaf boundaries S
    (halo keys — the tail of a rank's last leaf that lives on the next rank(s) — are fetched once
     per data set, not per build: the keys are immutable)
    all-reduce SUM   N x (ppm+2) x 8 B  leaf parameters, error bounds, key counts (zero where not owned)
The orchestration below is engine-agnostic: `CudaShardEngine` drives librmi_b200.so; the
CPU tests (gloo, world_size 2) plug in a numpy engine to exercise exactly this host logic.
"""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np
import torch
import torch.distributed as dist

from . import api

PHASE_TOP_LOCAL, PHASE_TOP_FINISH, PHASE_BOUNDS, PHASE_SPLIT, PHASE_LEAF, PHASE_STATS, PHASE_TOP_MID = range(7)
# Collectives of the top-model fit (include/rmi_b200.h rmi_shard_top_rounds): the first one follows
# TOP_LOCAL, the second one (two-round tops) follows TOP_MID.  "sum": all-reduce SUM of sums[0:8]
# as f64; "min": all-reduce MIN of sums[8:12] as signed 64-bit integers.
TOP_ROUNDS = {"linear_spline": (), "radix": (), "linear": ("sum",), "robust_linear": ("sum",),
              "normal": ("sum", "sum"), "lognormal": ("sum", "sum"), "cubic": ("min", "sum")}
# Table tops (rmi_shard_top_rounds == 4: one all-reduce MAX of the hint table / the pivots): offered by the one-call path
# (rmi_shard_train, NCCL) only — this host-driven orchestrator has no view of the library-owned table.
NATIVE_ONLY_TOPS = ("radix8", "radix18", "radix22", "radix26", "radix28", "histogram")
SHARDED_TOPS = tuple(TOP_ROUNDS) + NATIVE_ONLY_TOPS
_PPM = {"linear": 2, "robust_linear": 2, "linear_spline": 2, "loglinear": 2, "cubic": 4, "normal": 3, "lognormal": 3}
_TORCH_OF_KEY = {api.KEY_U64: torch.int64, api.KEY_U32: torch.int32, api.KEY_F64: torch.float64}


class _Ends(C.Structure):
    _fields_ = [("first_key_bits", C.c_uint64), ("last_key_bits", C.c_uint64), ("last_run_start", C.c_uint64),
                ("n_local", C.c_uint64), ("no_dups", C.c_uint64)]


class _Info(C.Structure):
    _fields_ = [("base", C.c_uint64), ("n_global", C.c_uint64), ("has_prev", C.c_int32), ("is_last", C.c_int32),
                ("prev_key_bits", C.c_uint64), ("prev_F", C.c_uint64), ("first_key_bits", C.c_uint64),
                ("last_key_bits", C.c_uint64), ("last_F", C.c_uint64), ("halo_capacity", C.c_uint64),
                ("no_dups", C.c_uint64), ("pivot_x", C.c_double), ("pivot_y", C.c_double)]


class _Buffers(C.Structure):
    _fields_ = [("sums", C.c_void_p), ("S", C.c_void_p), ("params", C.c_void_p), ("errors", C.c_void_p),
                ("counts", C.c_void_p), ("status", C.c_void_p)]


def key_bits_to_float(bits: int, key_type: int) -> float:
    """key.as_float() of a raw key (only used to place the common pivot of the sums)."""
    if key_type == api.KEY_F64:
        return struct.unpack("<d", struct.pack("<Q", bits))[0]
    return float(bits)


def plan_global_layout(ends_all: np.ndarray, key_type: int, num_leaves: int) -> list[dict]:
    """From every rank's (first_key_bits, last_key_bits, last_run_start, n_local) derive, for every
    rank, its shard description.  Pure function of the gathered table: every rank computes the same.

    prev_key / prev_F: last key before the slab and the first global index of its run of equal keys
    (the offset FixDupsIter would report, reference models/mod.rs:154-185), which may lie several
    ranks back when whole slabs consist of one repeated key."""
    world = ends_all.shape[0]
    n_local = [int(x) for x in ends_all[:, 3]]
    bases = [0]
    for g in range(world):
        bases.append(bases[-1] + n_local[g])
    n_global = bases[-1]
    nonempty = [g for g in range(world) if n_local[g] > 0]
    last_F = {}
    prev = None
    for g in nonempty:
        first_b, last_b, lrs = int(ends_all[g, 0]), int(ends_all[g, 1]), int(ends_all[g, 2])
        if lrs == 0 and prev is not None and int(ends_all[prev, 1]) == first_b:
            last_F[g] = last_F[prev]           # the whole slab is one run that began on an earlier rank
        else:
            last_F[g] = bases[g] + lrs
        prev = g
    # no two equal keys anywhere: every rank is duplicate-free and no cut separates two equal keys
    no_dups = ends_all.shape[1] > 4 and all(int(ends_all[g, 4]) == 1 for g in nonempty)
    for a, b in zip(nonempty, nonempty[1:]):
        if int(ends_all[a, 1]) == int(ends_all[b, 0]):
            no_dups = False
    first_bits = int(ends_all[nonempty[0], 0]) if nonempty else 0
    last_bits = int(ends_all[nonempty[-1], 1]) if nonempty else 0
    gl_last_F = last_F[nonempty[-1]] if nonempty else 0
    px = 0.5 * key_bits_to_float(first_bits, key_type) + 0.5 * key_bits_to_float(last_bits, key_type)
    py = 0.5 * float(num_leaves)
    out = []
    for g in range(world):
        before = [r for r in nonempty if r < g]
        p = before[-1] if before else None
        out.append(dict(base=bases[g], n_global=n_global, has_prev=int(p is not None),
                        is_last=int(bool(nonempty) and g == nonempty[-1]),
                        prev_key_bits=int(ends_all[p, 1]) if p is not None else 0,
                        prev_F=last_F[p] if p is not None else 0,
                        first_key_bits=first_bits, last_key_bits=last_bits, last_F=gl_last_F,
                        no_dups=int(bool(no_dups)), pivot_x=px, pivot_y=py, bases=bases))
    return out


def plan_halo(bases: list[int], v: list[int], n_global: int) -> list[tuple[int, int, int, int]]:
    """v[g] = first leaf boundary S[j] >= bases[g+1] (the end of rank g's last owned leaf).
    Rank g needs global keys [bases[g+1], min(v[g] + 1, n_global)); returns the transfers
    (dst_rank, src_rank, src_local_offset, count) that deliver them."""
    world = len(bases) - 1
    moves = []
    for g in range(world - 1):
        lo, hi = bases[g + 1], min(v[g] + 1, n_global)
        for r in range(g + 1, world):
            a, b = max(lo, bases[r]), min(hi, bases[r + 1])
            if b > a:
                moves.append((g, r, a - bases[r], b - a))
    return moves


_NP_OF_KEY = {api.KEY_U64: np.uint64, api.KEY_U32: np.uint32, api.KEY_F64: np.float64}


def key_type_of_path(path: str) -> int:
    """src/main.rs:122-132: the key type is taken from the file NAME."""
    import os
    name = os.path.basename(path)
    if "uint64" in name:
        return api.KEY_U64
    if "uint32" in name:
        return api.KEY_U32
    if "f64" in name:
        return api.KEY_F64
    raise api.RMIPanic("Data file must contain uint64, uint32, or f64.")


def slab_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Rank r's slab of an n-key array: [n*r/world, n*(r+1)/world) — contiguous, disjoint, covering."""
    return n * rank // world, n * (rank + 1) // world


def read_slab(path: str, rank: int, world: int, key_type: int | None = None) -> tuple[np.ndarray, int]:
    """Reads ONLY this rank's slab of a reference-format key file (u64 LE count + packed keys,
    README.md:26-31 / src/load.rs:132-157) into host memory; returns (keys, total key count).
    Every rank reads 1/world of the file, so a node's ranks load the data set in parallel."""
    kt = key_type_of_path(path) if key_type is None else key_type
    dt = np.dtype(_NP_OF_KEY[kt]).newbyteorder("<")
    with open(path, "rb") as f:
        head = f.read(8)
        if len(head) != 8:
            raise api.RMIPanic(f"Unable to read the key count of {path}")
        n = int(np.frombuffer(head, dtype="<u8")[0])
        lo, hi = slab_bounds(n, rank, world)
        f.seek(8 + lo * dt.itemsize)
        keys = np.fromfile(f, dtype=dt, count=hi - lo)
    if keys.size != hi - lo:
        raise api.RMIPanic(f"{path} is shorter than its header says ({n} keys)")
    return keys.astype(_NP_OF_KEY[kt], copy=False), n


class ShardedTrainingData:
    """This rank's slab of a globally sorted key array, in device memory with room behind it
    for halo keys.  `keys` must be a 1-D torch tensor on the rank's device (int64 storage for
    uint64 keys, int32 for uint32, float64)."""

    def __init__(self, keys: torch.Tensor, n_local: int | None = None, key_type: int = api.KEY_U64,
                 halo_capacity: int = 1 << 20, group=None):
        self.group = group
        self.key_type = key_type
        n_local = keys.numel() if n_local is None else n_local
        cap = n_local + halo_capacity
        if keys.numel() >= cap:
            self.buf = keys
        else:
            self.buf = torch.empty(cap, dtype=keys.dtype, device=keys.device)
            self.buf[:n_local].copy_(keys[:n_local])
        self.n_local = n_local
        self.halo_capacity = self.buf.numel() - n_local
        self.engine = CudaShardEngine(self)

    @classmethod
    def from_file(cls, path: str, device: torch.device, key_type: int | None = None, halo_capacity: int = 1 << 20,
                  group=None) -> "ShardedTrainingData":
        """The loader of a range-partitioned build: this rank's slab of the key file -> pinned host
        memory -> its GPU (src/load.rs:132-157 for one slab)."""
        rank, world = _world(group)
        kt = key_type_of_path(path) if key_type is None else key_type
        keys, _ = read_slab(path, rank, world, kt)
        host = torch.from_numpy(keys.view(np.int64) if kt == api.KEY_U64 else (keys.view(np.int32) if kt == api.KEY_U32 else keys))
        if torch.cuda.is_available():
            host = host.pin_memory()
        dev_keys = torch.empty(host.numel() + halo_capacity, dtype=host.dtype, device=device)
        dev_keys[: host.numel()].copy_(host, non_blocking=False)
        return cls(dev_keys, host.numel(), kt, halo_capacity, group)

    def grow_halo(self, capacity: int):
        """Re-home the slab in a buffer with room for `capacity` halo keys (a leaf reached further
        into the following ranks than expected — heavy skew)."""
        old = self.buf
        self.buf = torch.empty(self.n_local + capacity, dtype=old.dtype, device=old.device)
        self.buf[: self.n_local].copy_(old[: self.n_local])
        self.halo_capacity = capacity
        self.engine.end()
        self.engine = CudaShardEngine(self)
        for attr in ("_min_cap", "_halo_have"):
            if hasattr(self, attr):
                delattr(self, attr)


class CudaShardEngine:
    """Phases of a range-partitioned build on librmi_b200.so (include/rmi_b200.h rmi_shard_*)."""

    def __init__(self, data: ShardedTrainingData):
        self.data = data
        self.lib = api.load_library()
        L = self.lib
        L.rmi_shard_ends_get.argtypes = [C.c_void_p, C.POINTER(_Ends)]
        L.rmi_shard_build_create.argtypes = [C.c_void_p, C.POINTER(_Info), C.c_char_p, C.c_uint64, C.POINTER(_Buffers),
                                             C.c_void_p, C.POINTER(C.c_void_p)]
        L.rmi_shard_phase.argtypes = [C.c_void_p, C.c_int]
        L.rmi_shard_set_halo.argtypes = [C.c_void_p, C.c_uint64]
        L.rmi_shard_finish.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.POINTER(api._Result))]
        L.rmi_shard_build_destroy.argtypes = [C.c_void_p]
        L.rmi_shard_comm_unique_id.argtypes = [C.c_void_p]
        L.rmi_shard_comm_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.rmi_shard_comm_destroy.argtypes = [C.c_void_p]
        L.rmi_shard_set_partition.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int, C.c_int]
        L.rmi_shard_train.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.POINTER(api._Result))]
        self.device = data.buf.device
        self.ds = api.RMITrainingData.from_device(data.buf.data_ptr(), data.n_local, data.key_type,
                                                  self.device.index or 0, keep_alive=data.buf)
        self._build = None

    def ends(self):
        e = _Ends()
        api._check(self.lib.rmi_shard_ends_get(self.ds._h, C.byref(e)))
        return int(e.first_key_bits), int(e.last_key_bits), int(e.last_run_start), int(e.n_local), int(e.no_dups)

    def begin(self, info: dict, spec: str, num_leaves: int, bufs: dict):
        key = (spec, num_leaves, tuple(bufs[k].data_ptr() for k in sorted(bufs)))
        if self._build is not None and getattr(self, "_build_key", None) == key:
            return      # same spec / buffers: the build object (scratch, events) is reused
        self.end()
        self._build_key = key
        ci = _Info(base=info["base"], n_global=info["n_global"], has_prev=info["has_prev"], is_last=info["is_last"],
                   prev_key_bits=info["prev_key_bits"], prev_F=info["prev_F"], first_key_bits=info["first_key_bits"],
                   last_key_bits=info["last_key_bits"], last_F=info["last_F"], halo_capacity=self.data.halo_capacity,
                   no_dups=info.get("no_dups", 0), pivot_x=info["pivot_x"], pivot_y=info["pivot_y"])
        cb = _Buffers(*(bufs[k].data_ptr() for k in ("sums", "S", "params", "errors", "counts", "status")))
        h = C.c_void_p()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        api._check(self.lib.rmi_shard_build_create(self.ds._h, C.byref(ci), spec.encode(), num_leaves, C.byref(cb),
                                                   C.c_void_p(stream), C.byref(h)))
        self._build = h
        self._spec = spec

    def phase(self, k: int):
        api._check(self.lib.rmi_shard_phase(self._build, k))

    def set_halo(self, count: int):
        api._check(self.lib.rmi_shard_set_halo(self._build, count))

    def halo_view(self, offset: int, count: int) -> torch.Tensor:
        n = self.data.n_local
        return self.data.buf[n + offset: n + offset + count]

    def local_view(self, offset: int, count: int) -> torch.Tensor:
        return self.data.buf[offset: offset + count]

    def set_partition(self, bases: list[int], world: int, rank: int):
        arr = (C.c_uint64 * (world + 1))(*bases)
        api._check(self.lib.rmi_shard_set_partition(self._build, arr, world, rank))

    def train(self, comm, flags: int = 0):
        """The whole build in one library call: phases and NCCL collectives on the build's stream (rmi_shard_train)."""
        res = C.POINTER(api._Result)()
        api._check(self.lib.rmi_shard_train(self._build, comm, int(flags), C.byref(res)))
        return api.result_from_pointer(res, self._spec)

    def finish(self, flags: int = 0):
        res = C.POINTER(api._Result)()
        api._check(self.lib.rmi_shard_finish(self._build, int(flags), C.byref(res)))
        return api.result_from_pointer(res, self._spec)

    def end(self):
        if self._build is not None:
            self.lib.rmi_shard_build_destroy(self._build)
            self._build = None

    def __del__(self):
        try:
            self.end()
        except Exception:
            pass


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


_native_comms = {}


def native_comm(group, device: torch.device, single_rank_ok: bool = False):
    """The library's own NCCL communicator for `group` (rmi_shard_comm_*): rank 0 draws the unique id, the
    128 bytes travel through torch.distributed, every rank joins.  Cached per (group, device).  Returns None
    when the group does not run over NCCL (gloo tests; a single rank unless single_rank_ok — a one-rank
    communicator is how the one-call path is exercised on a one-GPU box) or NCCL cannot be loaded."""
    rank, world = _world(group)
    if device.type != "cuda" or (world > 1 and dist.get_backend(group) != "nccl") or (world <= 1 and not single_rank_ok):
        return None
    key = (id(group) if group is not None else 0, device.index or 0, world)
    if key in _native_comms:
        return _native_comms[key]
    lib = api.load_library()
    lib.rmi_shard_comm_unique_id.argtypes = [C.c_void_p]
    lib.rmi_shard_comm_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    buf = (C.c_uint8 * 128)()
    ok = 1
    if rank == 0:
        ok = 1 if lib.rmi_shard_comm_unique_id(buf) == 0 else 0
    if world > 1:
        t = torch.tensor(list(buf) + [ok], dtype=torch.uint8, device=device)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        vals = t.cpu().tolist()
    else:
        vals = list(buf) + [ok]
    comm = None
    if vals[128] == 1:
        ident = (C.c_uint8 * 128)(*vals[:128])
        h = C.c_void_p()
        api._check(lib.rmi_shard_comm_create(ident, world, rank, device.index or 0, C.byref(h)))
        comm = h
    _native_comms[key] = comm
    return comm


def train_sharded(data, model_spec: str, num_leaves: int, flags: int = 0, group=None, engine=None, counts: bool = True,
                  native: bool | None = None):
    """rmi_lib::train on a range-partitioned key array; returns the full TrainedRMI on every rank.

    native=None (default): with the CUDA engine over an NCCL group the whole build is ONE library call
    (rmi_shard_train: kernels and collectives on one stream, leaf records exchanged as an all-gather by
    ownership range); otherwise — gloo, the numpy engine of the CPU tests, native=False — the phases are
    sequenced here and the collectives go through torch.distributed."""
    eng = engine if engine is not None else data.engine
    group = group if group is not None else getattr(data, "group", None)
    rank, world = _world(group)
    dev = eng.device
    parts = model_spec.split(",")
    if len(parts) != 2:
        raise api.RMIPanic("only two-layer RMIs can be trained (the reference panics on other depths)")
    if parts[0] not in SHARDED_TOPS:
        raise api.RMIError(f"range-partitioned builds offer the top models {SHARDED_TOPS}")
    ppm = _PPM.get(parts[1])
    if ppm is None:
        raise api.RMIPanic(f"unsupported or unknown leaf model {parts[1]}")
    N = int(num_leaves)

    # 1. what every rank's slab looks like at its ends (cached on the data object: the data is immutable)
    layout = getattr(data, "_layout", None)
    if layout is None or layout[0] != N:
        e = torch.tensor(np.array(eng.ends(), dtype=np.uint64).view(np.int64), dtype=torch.int64, device=dev)
        gathered = [torch.empty_like(e) for _ in range(world)]
        if world > 1:
            dist.all_gather(gathered, e, group=group)
        else:
            gathered = [e]
        ends_all = torch.stack(gathered).cpu().numpy().view(np.uint64)
        layout = (N, plan_global_layout(ends_all, data.key_type, N))
        data._layout = layout
    info = layout[1][rank]
    bases, n_global = info["bases"], info["n_global"]

    bufs = getattr(data, "_bufs", None)
    if bufs is None or bufs["S"].numel() != N + 1 or bufs["params"].numel() != N * ppm:
        # params | errors | counts live in ONE allocation so that a single all-reduce combines them
        rec = torch.empty(N * (ppm + 2), dtype=torch.int64, device=dev)
        bufs = dict(sums=torch.zeros(16, dtype=torch.float64, device=dev),   # [0:8] f64 sums, [8:16] i64 slots
                    S=torch.empty(N + 1, dtype=torch.int64, device=dev),
                    params=rec[: N * ppm].view(torch.float64),
                    errors=rec[N * ppm: N * (ppm + 1)],
                    counts=rec[N * (ppm + 1):],
                    status=torch.zeros(1, dtype=torch.int32, device=dev), records=rec)
        data._bufs = bufs
    eng.begin(info, model_spec, N, bufs)

    comm = None
    if native is not False and isinstance(eng, CudaShardEngine):
        comm = native_comm(group, dev, single_rank_ok=native is True)
        if native is True and comm is None:
            raise api.RMIError("native=True needs an NCCL process group (or a single rank) and a loadable libnccl.so.2")
    if comm is not None:
        # halo keys are fetched once per data set (see step 4 below), then the whole build is one call
        if getattr(data, "_halo_have", None) is None:
            if world > 1:
                cap = _min_halo_capacity(data, group, world, dev)
                moves = plan_halo(bases, [bases[g + 1] + cap - 1 for g in range(world)], n_global)
                data._halo_have = _exchange_halo(eng, moves, rank, group, dev)
            else:
                data._halo_have = 0
        eng.set_halo(data._halo_have or 0)
        eng.set_partition(bases, world, rank)
        try:
            return eng.train(comm, int(flags) | (api.FLAG_LEAF_COUNTS if counts else 0))
        except api.RMIPanic as e:
            if "halo" not in str(e):
                raise
        return _retry_with_larger_halo(data, bufs, bases, n_global, N, model_spec, num_leaves, flags, group, world, dev, counts, native)

    if parts[0] in NATIVE_ONLY_TOPS:
        raise api.RMIError(f"a range-partitioned build with the top model {parts[0]} needs the one-call path "
                           "(rmi_shard_train over an NCCL process group, or native=True with a single rank)")
    # 2. top model: local part -> tiny all-reduce(s) -> closed form (identical on every rank)
    def top_collective(kind):
        if world <= 1:
            return
        if kind == "sum":
            dist.all_reduce(bufs["sums"][:8], op=dist.ReduceOp.SUM, group=group)
        else:
            dist.all_reduce(bufs["sums"].view(torch.int64)[8:12], op=dist.ReduceOp.MIN, group=group)

    rounds = TOP_ROUNDS[parts[0]]
    eng.phase(PHASE_TOP_LOCAL)
    if rounds:
        top_collective(rounds[0])
    if len(rounds) > 1:
        eng.phase(PHASE_TOP_MID)
        top_collective(rounds[1])
    eng.phase(PHASE_TOP_FINISH)
    # 3. leaf boundaries: local lower bounds -> all-reduce MIN
    eng.phase(PHASE_BOUNDS)
    if world > 1:
        dist.all_reduce(bufs["S"], op=dist.ReduceOp.MIN, group=group)
    eng.phase(PHASE_SPLIT)
    # 4. halo: the keys of a rank's last leaf that live on the following rank(s).  The keys are
    #    immutable, so the first `halo_capacity` keys behind every slab are fetched ONCE per data set
    #    (no planning, no host synchronisation, no transfer inside a build); a build whose last leaf
    #    reaches further reports it through the status word and takes the planned path below.
    if world > 1 and getattr(data, "_halo_have", None) is None:
        cap = _min_halo_capacity(data, group, world, dev)
        moves = plan_halo(bases, [bases[g + 1] + cap - 1 for g in range(world)], n_global)
        data._halo_have = _exchange_halo(eng, moves, rank, group, dev)
    eng.set_halo(getattr(data, "_halo_have", 0) or 0)
    # 5. leaves owned by this rank, then everyone gets everything
    eng.phase(PHASE_LEAF)
    if world > 1:
        rec = bufs["records"]
        dist.all_reduce(rec if counts else rec[: N * (ppm + 1)], op=dist.ReduceOp.SUM, group=group)
        # the status word is a BIT MASK (kernels.h StatusBit): combine by OR, not MAX, so that every rank sees every
        # rank's bits and all of them take the same decision below (retry with a larger halo / raise)
        st_all = [torch.empty_like(bufs["status"]) for _ in range(world)]
        dist.all_gather(st_all, bufs["status"], group=group)
        acc = st_all[0].clone()
        for t in st_all[1:]:
            acc |= t
        bufs["status"].copy_(acc)
    eng.phase(PHASE_STATS)
    try:
        return eng.finish(int(flags) | (api.FLAG_LEAF_COUNTS if counts else 0))
    except api.RMIPanic as e:
        if world <= 1 or "halo" not in str(e):
            raise
    return _retry_with_larger_halo(data, bufs, bases, n_global, N, model_spec, num_leaves, flags, group, world, dev, counts, native)


def _retry_with_larger_halo(data, bufs, bases, n_global, N, model_spec, num_leaves, flags, group, world, dev, counts, native):
    # A leaf reaches past the prefetched halo (heavy skew).  The status word is the OR over ranks, so
    # every rank arrives here together: size the halo from the global boundaries S and build again.
    S = bufs["S"]
    cuts = torch.tensor(bases[1:], dtype=torch.int64, device=dev)
    pos = torch.searchsorted(S, cuts).clamp_(max=N)
    v = S[pos].cpu().tolist()
    need = {}
    for (dst, src, off, cnt) in plan_halo(bases, v, n_global):
        need[dst] = need.get(dst, 0) + cnt
    most = max(need.values(), default=0)
    if most <= _min_halo_capacity(data, group, world, dev) or not hasattr(data, "grow_halo"):
        raise api.RMIError(f"a leaf reaches further into the next rank than the halo capacity ({most} keys needed)")
    data.grow_halo(int(most * 1.25) + 1024)
    return train_sharded(data, model_spec, num_leaves, flags, group, counts=counts, native=native)


def _exchange_halo(eng, moves, rank, group, dev) -> int:
    """Carries out plan_halo's transfers; returns how many halo keys this rank now holds."""
    # gloo cannot send/recv device memory (one-GPU test boxes): stage through the host there
    stage = dev.type == "cuda" and dist.get_backend(group) == "gloo"
    ops, recv_off, landed = [], 0, []
    for (dst, src, off, cnt) in moves:
        if dst == rank:
            view = eng.halo_view(recv_off, cnt)
            t = torch.empty(cnt, dtype=view.dtype) if stage else view
            if stage:
                landed.append((view, t))
            ops.append(dist.P2POp(dist.irecv, t, src, group=group))
            recv_off += cnt
        elif src == rank:
            view = eng.local_view(off, cnt)
            ops.append(dist.P2POp(dist.isend, view.cpu() if stage else view, dst, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for view, t in landed:
        view.copy_(t)
    return recv_off


def _min_halo_capacity(data, group, world, dev):
    cap = getattr(data, "_min_cap", None)
    if cap is None:
        t = torch.tensor([data.halo_capacity], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        cap = int(t.item())
        data._min_cap = cap
    return cap
