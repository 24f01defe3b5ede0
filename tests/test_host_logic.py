"""CPU tests of the host-side logic: C-ABI symbol export, metric closed forms vs sklearn, subject
factorisation, statistic broadcasting, batch sharding, world_size-2 gloo gradient exchange."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from multimodal_error_detection_b200 import build
    return build.build()


def test_cabi_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "b200med.h")).read()
    declared = set(re.findall(r"\b(b200med_[a-z0-9_]+)\s*\(", header))
    declared.discard("b200med_stream_desc")
    lib = ctypes.CDLL(built_lib)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    from multimodal_error_detection_b200 import _lib
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().b200med_version() == 100


def test_no_cpu_fallback(built_lib):
    """Without a GPU every compute entry point must fail loudly."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from multimodal_error_detection_b200.dataset import dataset_utils
    from multimodal_error_detection_b200.modeling import models
    with pytest.raises(RuntimeError):
        dataset_utils.powerset_error_labels(torch.zeros(4, 5))
    fe = models.FeatureExtractor(64, 8, [16])
    with pytest.raises(RuntimeError):
        fe(torch.zeros(2, 64))


def test_tail_shape_rule_and_argument_errors(built_lib):
    """Host logic of csrc/mlp_tail.cu and the multi-copy entry point, no kernel launched: which layer shapes the fused head tail
    serves (weights + one 64-row tile within the shared memory of an SM, hidden widths 64 / 128 / 256, <= 8 output columns), the
    slab count of the statistics partials, and the C-ABI error convention (B200MED_E_ARG = -1 with a message before any CUDA
    call) for unsupported shapes, misaligned pointers and malformed pointer tables."""
    import ctypes as C
    from multimodal_error_detection_b200 import _lib
    lib = _lib.load()
    sup = lib.b200med_tail_supported
    # the LSTM head: 128 -> 256 -> 64 -> C, forward and backward kinds
    assert sup(128, 256, 0) and sup(256, 64, 0) and sup(64, 1, 2) and sup(64, 8, 2)
    assert sup(64, 256, 3) and sup(256, 128, 1)
    assert not sup(64, 9, 2)                      # more than 8 output columns
    assert not sup(256, 32, 0) and not sup(32, 16, 0)      # the CNN head's narrow layers run layer by layer
    assert not sup(96, 256, 3)                    # the last hidden width must be 64 / 128 / 256
    assert not sup(130, 256, 0)                   # K must be a multiple of 4
    assert not sup(512, 256, 0) and not sup(512, 256, 1)   # over the shared memory of an SM
    assert [lib.b200med_tail_slabs(m) for m in (1, 64, 65, 8192)] == [1, 1, 2, 128]
    null = C.c_void_p(0)
    one = C.c_void_p(16)
    rc = lib.b200med_tail_fwd_hidden(one, 8, 256, 0, 0, null, null, null, 1e-5, 0.1, null, null, null, null, null, null, one, null, 32,
                                     one, null, null)
    assert rc == -1 and b"unsupported layer shape" in lib.b200med_last_error()
    rc = lib.b200med_tail_fwd_hidden(C.c_void_p(4), 8, 128, 0, 0, null, null, null, 1e-5, 0.1, null, null, null, null, null, null, one,
                                     null, 256, one, null, null)
    assert rc == -1 and b"16-byte aligned" in lib.b200med_last_error()
    rc = lib.b200med_tail_fwd_out(one, 8, 64, 0, 1, null, null, null, 1e-5, 0.1, null, null, null, null, null, null, one, null, 1, one,
                                  null)
    assert rc == -1 and b"batch statistics need" in lib.b200med_last_error()
    rc = lib.b200med_tail_bwd_out(one, 9, one, one, 8, 64, one, one, one, null)
    assert rc == -1
    rc = lib.b200med_multi_copy_f32(null, null, null, 3, 0, null, 0.9, 0.999, null)
    assert rc == -1 and b"bad arguments" in lib.b200med_last_error()
    rc = lib.b200med_multi_copy_f32(null, null, null, 0, 7, null, 0.9, 0.999, null)
    assert rc == -1 and b"dst_dtype" in lib.b200med_last_error()
    assert lib.b200med_multi_copy_f32(null, null, null, 0, 0, null, 0.9, 0.999, null) == 0      # nothing to do


def test_fused_tail_is_chosen_for_the_lstm_head_only(built_lib):
    """heads._fused_ok (host logic): the three-kernel tail takes the LSTM head's 128 -> 256 -> 64 -> C layers (MED/modeling/
    models.py:166-186) and leaves the CNN head's 256 -> 32 -> 16 -> C tail (models.py:100-110) to the layer-at-a-time kernels;
    the switch B200MED_FUSED_TAIL / heads.FUSED_TAIL turns it off for A/B timing."""
    import torch.nn as nn
    from multimodal_error_detection_b200 import heads
    from multimodal_error_detection_b200.modeling.models import CNN, LSTM

    def tensors(seq):
        t = []
        for m in seq:
            if isinstance(m, nn.Linear):
                t += [m.weight, m.bias]
            elif isinstance(m, nn.BatchNorm1d):
                t += [m.weight, m.bias, m.running_mean, m.running_var, m.num_batches_tracked]
        return t

    for C_ in (1, 6):
        lstm = LSTM(58, 16, 3, 128, C_)
        assert heads._fused_ok(torch.zeros(8, 128), tensors(lstm.linear_layers), 2)
    cnn = CNN(58, 10, 1)
    assert not heads._fused_ok(torch.zeros(8, 128), tensors(cnn.linear_layers), 3)
    lstm = LSTM(58, 16, 3, 128, 1)
    assert not heads._fused_ok(torch.zeros(8, 96), tensors(lstm.linear_layers), 2)        # input width != the first layer's
    prev, heads.FUSED_TAIL = heads.FUSED_TAIL, False
    try:
        assert not heads._fused_ok(torch.zeros(8, 128), tensors(lstm.linear_layers), 2)
    finally:
        heads.FUSED_TAIL = prev


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "multimodal_error_detection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


@pytest.mark.parametrize("C", [2, 6])
def test_metrics_match_sklearn(C):
    from sklearn.metrics import accuracy_score, confusion_matrix, f1_score, jaccard_score
    from multimodal_error_detection_b200 import metrics as M
    rng = np.random.Generator(np.random.PCG64(C))
    for trial in range(30):
        n = int(rng.integers(1, 80))
        hi = C if trial % 3 else max(1, C - 2)          # sometimes leave classes absent
        y, p = rng.integers(0, hi, n), rng.integers(0, hi, n)
        cm = confusion_matrix(y, p, labels=list(range(C)))
        assert abs(M.accuracy(cm) - accuracy_score(y, p)) < 1e-12
        for avg in ("macro", "weighted"):
            assert abs(M.f1_avg(cm, avg) - f1_score(y, p, average=avg, zero_division=0)) < 1e-12
            assert abs(M.jaccard_avg(cm, avg) - jaccard_score(y, p, average=avg, zero_division=0)) < 1e-12
        assert np.array_equal(M.sklearn_cm(cm), confusion_matrix(y, p))
        if C == 2 and len(set(y) | set(p)) == 2:
            assert abs(M.f1_binary(cm) - f1_score(y, p, average="binary", zero_division=0)) < 1e-12
            assert abs(M.jaccard_binary(cm) - jaccard_score(y, p, average="binary", zero_division=0)) < 1e-12
        if C == 6:
            yb, pb = (y != 0).astype(int), (p != 0).astype(int)
            assert np.array_equal(M.binarise(cm), confusion_matrix(yb, pb, labels=[0, 1]))


def test_factorize_subjects_appearance_order():
    from multimodal_error_detection_b200.table import factorize_subjects
    codes, uniq = factorize_subjects(["z", "z", "a", "m", "a", "z"])
    assert uniq == ["z", "a", "m"] and codes.tolist() == [0, 0, 1, 2, 1, 0]


def test_expand_stat_shapes():
    from multimodal_error_detection_b200.ops import expand_stat
    D, W = 6, 4
    for stat in (2.0, torch.arange(D).float(), torch.arange(D).float().reshape(1, D)):
        assert expand_stat(stat, D, W, "cpu").shape == (1, D)
    per_step = torch.arange(W * D).float().reshape(W, D)
    assert torch.equal(expand_stat(per_step, D, W, "cpu"), per_step)


def test_shard_batch_partitions():
    from multimodal_error_detection_b200.parallel import shard_batch
    idx = torch.arange(13)
    for world in (1, 2, 4, 8):
        parts = [shard_batch(idx, r, world) for r in range(world)]
        assert torch.equal(torch.cat(parts), idx)
        sizes = [p.numel() for p in parts]
        assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1          # even split: no rank is left without a step


def test_sampler_matches_torch_dataloader():
    """Batch composition = a stock DataLoader with the reference's generator recipe (dataset_utils.py:526)."""
    from torch.utils.data import DataLoader, TensorDataset
    from multimodal_error_detection_b200.dataset.CustomWindowDataset import _IndexOnly
    n, B = 37, 8
    ref = DataLoader(TensorDataset(torch.arange(n)), batch_size=B, shuffle=True, generator=torch.Generator().manual_seed(42))
    mine = DataLoader(_IndexOnly(n), batch_size=B, shuffle=True, generator=torch.Generator().manual_seed(42),
                      collate_fn=lambda items: torch.as_tensor(items))
    for _ in range(3):  # generator state carries across epochs in both
        assert [b[0].tolist() for b in ref] == [b.tolist() for b in mine]


def test_index_batches_mirror_the_dataloader_exactly():
    """DeviceWindowLoader.index_batches() (vectorised: one randperm per pass) yields the batches of the stock DataLoader
    bit for bit, epoch after epoch, including the generator state left behind by a full pass, by an abandoned pass
    (max_batches) and in the unshuffled case; world_size 2 shards every global batch contiguously."""
    from torch.utils.data import DataLoader, TensorDataset
    from multimodal_error_detection_b200.dataset.CustomWindowDataset import DeviceWindowLoader

    class _DS:
        def __init__(self, n): self.n = n
        def __len__(self): return self.n

    for n, B in ((37, 8), (64, 16), (5, 8)):
        ref = DataLoader(TensorDataset(torch.arange(n)), batch_size=B, shuffle=True, generator=torch.Generator().manual_seed(42))
        mine = DeviceWindowLoader(_DS(n), B, shuffle=True, generator=torch.Generator().manual_seed(42))
        assert len(mine) == len(ref)
        for epoch in range(4):
            cap = 2 if epoch == 1 else None             # epoch 1 is abandoned after two batches in both
            mine.max_batches = cap
            want = []
            for k, b in enumerate(ref):
                if cap is not None and k >= cap:
                    break
                want.append(b[0].tolist())
            got = [b.tolist() for b in mine.index_batches()]
            assert got == want, (n, B, epoch)
    seq = DeviceWindowLoader(_DS(21), 8, shuffle=False)
    assert [b.tolist() for b in seq.index_batches()] == [list(range(0, 8)), list(range(8, 16)), list(range(16, 21))]
    a = DeviceWindowLoader(_DS(37), 8, shuffle=True, generator=torch.Generator().manual_seed(42))
    r0 = DeviceWindowLoader(_DS(37), 8, shuffle=True, generator=torch.Generator().manual_seed(42), rank=0, world_size=2)
    r1 = DeviceWindowLoader(_DS(37), 8, shuffle=True, generator=torch.Generator().manual_seed(42), rank=1, world_size=2)
    for g, x, y in zip(a.index_batches(), r0.index_batches(), r1.index_batches()):
        assert torch.cat([x, y]).tolist() == g.tolist()
    # 8 ranks, 37 windows in batches of 16: the last global batch (5 windows < 8 ranks) is dropped on every rank, so all
    # ranks take the same number of steps (each step ends in a gradient all-reduce)
    per_rank = [[b.numel() for b in DeviceWindowLoader(_DS(37), 16, shuffle=True, generator=torch.Generator().manual_seed(42),
                                                      rank=r, world_size=8).index_batches()] for r in range(8)]
    assert all(len(p) == 2 and min(p) >= 1 for p in per_rank) and [sum(x) for x in zip(*per_rank)] == [16, 16]


def test_data_parallel_weights_of_uneven_shards():
    """A short last batch splits into shards that differ by one window: each rank's mean-reduced gradient is weighted by
    n_local * world / n_global so that the SUM all-reduce scaled by 1 / world is the mean over the GLOBAL batch."""
    from multimodal_error_detection_b200.dataset.CustomWindowDataset import DeviceWindowLoader

    class _DS:
        def __init__(self, n): self.n = n
        def __len__(self): return self.n

    world, n, B = 4, 37, 16                                   # global batches of 16, 16, 5 windows
    loaders = [DeviceWindowLoader(_DS(n), B, shuffle=True, generator=torch.Generator().manual_seed(42), rank=r, world_size=world)
               for r in range(world)]
    per_rank = [list(l.index_batches_with_global()) for l in loaders]
    assert [g for _, g in per_rank[0]] == [16, 16, 5]
    x = torch.arange(n, dtype=torch.float64) ** 1.5           # "per-window loss"
    for k in range(3):
        shards = [per_rank[r][k][0] for r in range(world)]
        n_global = per_rank[0][k][1]
        whole = torch.cat(shards)
        assert whole.numel() == n_global
        # sum over ranks of weight * local mean, scaled by 1 / world == global mean
        combined = sum(loaders[r].dp_weight(shards[r].numel(), n_global) * x[shards[r]].mean() for r in range(world)) / world
        assert abs(float(combined) - float(x[whole].mean())) < 1e-12
        if n_global == B:
            assert all(loaders[r].dp_weight(shards[r].numel(), n_global) == 1.0 for r in range(world))
    assert DeviceWindowLoader(_DS(n), B).dp_weight(5, 5) == 1.0


def test_lstm_dg_column_permutation():
    """Host side of the recurrence kernels' dG layout (csrc/lstm_rec.cu): column' = chunk*128 + unit_quarter*32 + gate*8 + i
    holds gate column gate*H + unit_quarter*32 + chunk*8 + i -- a permutation, and the two index tensors are inverses."""
    from multimodal_error_detection_b200.lstm_stack import _dg_perm
    H = 128
    orig_of, colp_of = _dg_perm(H, "cpu")
    assert sorted(orig_of.tolist()) == list(range(4 * H))
    assert torch.equal(orig_of[colp_of], torch.arange(4 * H)) and torch.equal(colp_of[orig_of], torch.arange(4 * H))
    for colp in (0, 7, 8, 31, 32, 127, 128, 300, 511):
        c, uq, g, i = colp // 128, (colp % 128) // 32, (colp % 32) // 8, colp % 8
        assert int(orig_of[colp]) == g * H + uq * 32 + c * 8 + i
    # un-permuting the rows of a permuted matrix restores it (what the backward does to dW and db)
    m = torch.arange(4 * H * 3, dtype=torch.float32).reshape(4 * H, 3)
    assert torch.equal(m.index_select(0, orig_of).index_select(0, colp_of), m)


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from multimodal_error_detection_b200 import parallel
rank, _, world = parallel.init_from_env("gloo")
torch.manual_seed(0)
w = torch.nn.Linear(16, 1)
x, y = torch.randn(12, 16), (torch.rand(12) > 0.5).float()
def grads(rows):
    w.zero_grad()
    # SUM-reduced per-rank loss, scaled by the GLOBAL batch: what the all-reduce + 1/world Adam scale computes
    loss = torch.nn.functional.binary_cross_entropy_with_logits(w(x[rows]).squeeze(1), y[rows], reduction="sum") / 12
    loss.backward()
    return torch.cat([p.grad.reshape(-1) for p in w.parameters()])
full = grads(torch.arange(12))
mine = grads(parallel.shard_batch(torch.arange(12), rank, world)).clone()
parallel.allreduce_sum_(mine)
assert torch.allclose(mine, full, atol=1e-6), (mine, full)
t = parallel.max_over_ranks(float(rank + 1), device="cpu")
assert t == world
dist.destroy_process_group()
print("ok", rank)
'''


def test_gloo_world2_gradient_exchange(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29731",
                   CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


_PEER_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from multimodal_error_detection_b200 import parallel
rank, _, world = parallel.init_from_env("gloo")
# no CUDA device here: the allocation of the peer buffer fails on every rank; the set-up must notice that COLLECTIVELY (every
# rank raises, nobody is left waiting in a collective) so that the callers fall back to the NCCL / gloo all-reduce together
try:
    parallel.PeerAllReduce(1000, "cpu")
    raise SystemExit("PeerAllReduce succeeded without a GPU")
except RuntimeError as e:
    assert "peer-memory exchange unavailable" in str(e), str(e)
# the ranks are still in step: a collective after the failed set-up completes
t = torch.tensor([float(rank + 1)])
dist.all_reduce(t)
assert float(t) == 3.0
dist.destroy_process_group()
print("ok", rank)
'''


def test_gloo_world2_peer_exchange_setup_fails_collectively(tmp_path, built_lib):
    """parallel.PeerAllReduce (the NVLink peer-memory gradient exchange) negotiates its set-up: where the mappings cannot be
    made -- here: no GPU at all -- EVERY rank gets the RuntimeError and the process group stays usable (FusedAdam then keeps
    the all-reduce of torch.distributed, with a RuntimeWarning)."""
    script = tmp_path / "peer_worker.py"
    script.write_text(_PEER_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29733",
                   CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_split_k_choice_fills_whole_waves(built_lib):
    """b200med_gemm_bf16_pick_split (host logic, no GPU needed: 148 SMs assumed without a device): the weight-gradient shapes of
    the headline step get a split whose tiles x splits work items fill the SMs (or SM pairs) in whole waves -- the first
    heuristic gave the layer-1 gradient 160 items on 74 pairs = 3 waves at 72 %."""
    import ctypes as C
    lib = C.CDLL(built_lib)
    lib.b200med_gemm_bf16_pick_split.restype = C.c_int32
    lib.b200med_gemm_bf16_pick_split.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int32]
    rows = 8192 * 16
    for M, N, pair in ((512, 2048, True), (512, 256, True), (256, 512, True), (512, 192, True), (32, 256, False)):      # N = 192: three quarters of a pair tile
        s = lib.b200med_gemm_bf16_pick_split(M, N, rows, 0)
        assert 1 <= s <= 160 and (rows // 64) // s >= 8, (M, N, s)
        bn = 256 if N >= 256 else 128 if N > 64 else 64
        tiles = (-(-M // 256)) * (-(-N // 256)) if pair else (-(-M // 128)) * (-(-N // bn))
        units = 74 if pair else 148
        items = tiles * s
        waves = -(-items // units)
        assert items / (waves * units) >= 0.85 or items <= units, (M, N, s, items, waves)
    assert lib.b200med_gemm_bf16_pick_split(512, 2048, rows, 0) == 9
    assert lib.b200med_gemm_bf16_pick_split(512, 2048, 512, 0) == 1          # short K: no split


def test_fp32_tensor_core_route_host_logic(built_lib):
    """ops._fp32_tc_split: one accumulation of the six-product GEMM never runs over more than 12 k-blocks (the tensor core
    truncates when it adds into the fp32 accumulator); ops._fp32_tc: without a GPU there is no such route (and no fallback:
    the fp32 ops themselves raise on CPU tensors)."""
    from multimodal_error_detection_b200 import ops
    assert ops._fp32_tc_split(6 * 2048) == 16 and ops._fp32_tc_split(6 * 512) == 4 and ops._fp32_tc_split(6 * 32) == 1
    for k6 in (64, 700, 6 * 58, 6 * 2048, 6 * 131072):
        s = ops._fp32_tc_split(k6)
        assert 1 <= s <= 256 and (s == 256 or -(-(-(-k6 // 64)) // s) <= ops.FP32_TC_KB_PER_SPLIT), (k6, s)
    if not torch.cuda.is_available():
        assert not ops._fp32_tc(32768, 512, 2048, 6 * 2048)
        with pytest.raises(RuntimeError):
            ops.linear_fwd_f32(torch.zeros(4, 8), torch.zeros(2, 8), None, relu=False)
    assert not ops._fp32_tc(64, 64, 64, 384)                    # small products stay on the fp32 FMA kernels
    assert not ops._fp32_tc(32768, 512, 58, 6 * 58)             # 6 x 58 elements per row: not a multiple of 16 bytes


def test_epoch_scores_equal_the_per_batch_scores():
    """modeling_utils._epoch_scores (all batches of an epoch at once) == the per-batch `_batch_scores` accumulated batch after
    batch, bit for bit: random counts plus the degenerate batches (one label present, empty predictions, empty batch)."""
    from multimodal_error_detection_b200.modeling import modeling_utils as mu
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 400, size=(64, 4))
    counts[3] = (50, 0, 0, 0)        # only label 0 present: sklearn's 1x1 confusion matrix is broadcast (reference quirk)
    counts[7] = (0, 0, 0, 9)         # only label 1
    counts[11] = (0, 12, 0, 0)       # y_true all 0, predictions all 1
    counts[12] = (0, 0, 5, 0)
    counts[20] = (0, 0, 0, 0)
    tot, cm = np.zeros(4), np.zeros((2, 2), dtype=int)
    for c in counts:
        f1, f1w, acc, jac, contrib = mu._batch_scores(c)
        tot += (f1, f1w, acc, jac)
        cm += contrib
    tot2, cm2 = mu._epoch_scores(counts)
    assert np.array_equal(tot, tot2), (tot, tot2)
    assert np.array_equal(cm, cm2), (cm, cm2)
    t0, c0 = mu._epoch_scores(np.zeros((0, 4), dtype=np.int64))
    assert not t0.any() and not c0.any()
