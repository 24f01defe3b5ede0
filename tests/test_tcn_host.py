"""CPU tests of the TeCNo autograd node's host logic: the kernels are replaced by their torch definitions
(tests/tcn_reference.py) and the node is compared with the oracle TeCNo (oracle/nets.py, pinned to the reference's
models_TCN.py by tests/golden) -- forward, input gradient and every parameter gradient, causal and non-causal,
one video and ragged batches."""
import numpy as np
import pytest
import torch

import tcn_reference as ref
from multimodal_error_detection_b200 import tcn
from multimodal_error_detection_b200.modeling import models_TCN
from oracle import nets


@pytest.fixture()
def emulated(monkeypatch):
    monkeypatch.setattr(tcn, "ops", ref.EmulatedOps)
    monkeypatch.setattr(tcn.StageConfig, "ptr_table", lambda self, layer_params: list(layer_params))
    monkeypatch.setattr(tcn, "_check", lambda x, params: None)
    monkeypatch.setattr(tcn, "require_cuda", lambda x: None)   # the emulation runs the same code path on host tensors


def _pair(causal, f_dim=58, stages=2, layers=4):
    torch.manual_seed(0)
    o = nets.OracleTeCNo(stages, layers, 64, f_dim, 2, causal)
    m = models_TCN.MultiStageModel(stages, layers, 64, f_dim, 2, causal)
    m.load_state_dict(o.state_dict())
    nets.disable_dropout(o)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return o.train(), m.train()


@pytest.mark.parametrize("causal", [True, False])
def test_stage_node_matches_oracle(emulated, causal):
    o, m = _pair(causal)
    T = 77
    base = torch.randn(1, T, 58)
    xo = base.clone().requires_grad_(True)
    xm = base.clone().requires_grad_(True)
    out_o = o(xo.permute(0, 2, 1))
    out_m = m(xm.permute(0, 2, 1))
    assert m.impl == "b200" and out_m.shape == out_o.shape == (2, 1, 2, T)
    assert torch.allclose(out_m, out_o, rtol=1e-4, atol=1e-5)
    w = torch.randn_like(out_o)
    (out_o * w).sum().backward()
    (out_m * w).sum().backward()
    assert torch.allclose(xm.grad, xo.grad, rtol=1e-3, atol=1e-5)
    po, pm = dict(o.named_parameters()), dict(m.named_parameters())
    assert list(po) == list(pm)
    for k in po:
        scale = po[k].grad.abs().max().item() + 1e-12
        assert (pm[k].grad - po[k].grad).abs().max().item() <= 2e-4 * scale, k


def test_ragged_batch_equals_per_video(emulated):
    _, m = _pair(True, layers=5)
    m.eval()
    lengths = [40, 9, 65]
    frames = torch.randn(sum(lengths), 58)
    with torch.no_grad():
        cat = m.forward_ragged(frames, lengths)
        parts, s = [], 0
        for n in lengths:
            parts.append(m(frames[s:s + n].unsqueeze(0).permute(0, 2, 1))[:, 0])
            s += n
    assert torch.allclose(cat, torch.cat(parts, dim=2), rtol=1e-5, atol=1e-6)


def test_pack_and_grad_record_layouts():
    torch.manual_seed(1)
    wd, bd, w1, b1 = torch.randn(64, 64, 3), torch.randn(64), torch.randn(64, 64, 1), torch.randn(64)
    p = ref.pack_layer(wd, bd, w1, b1)
    assert p.numel() == ref.PACK == tcn.ops.TCN_PACK_FLOATS
    k, ci, co = 2, 5, 17
    assert p[ref.OFF_WDF + (k * 64 + ci) * 64 + co] == wd[co, ci, k]
    assert p[ref.OFF_WDB + (k * 64 + co) * 64 + ci] == wd[co, ci, k]
    assert p[ref.OFF_W1F + ci * 64 + co] == w1[co, ci, 0] and p[ref.OFF_W1B + co * 64 + ci] == w1[co, ci, 0]
    assert p[ref.OFF_BIAS + co] == bd[co] and p[ref.OFF_BIAS + 64 + co] == b1[co]
    assert ref.GRAD == tcn.ops.TCN_GRAD_FLOATS == 3 * 4096 + 4096 + 128


def test_ragged_geometry():
    tl, tr = tcn.ragged_geometry([3, 1, 2], "cpu")
    assert tl.tolist() == [0, 1, 2, 0, 0, 1] and tr.tolist() == [2, 1, 0, 0, 1, 0]


def test_unsupported_shape_and_cpu_raise():
    """No second implementation: a shape outside the fused kernels is a ValueError, a CPU tensor a RuntimeError."""
    m64 = models_TCN.MultiStageModel(2, 3, 64, 10, 2, True).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m64(torch.randn(1, 10, 20))
    m = models_TCN.MultiStageModel(2, 3, 32, 10, 2, True).eval()     # 32 feature maps: not the specialised shape
    with pytest.raises((ValueError, RuntimeError)):
        m(torch.randn(1, 10, 20))
    assert not m.stage1.fused_supported() and m64.stage1.fused_supported()


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from multimodal_error_detection_b200 import parallel
from multimodal_error_detection_b200.dataset.CustomFrameDataset import FrameLoader
rank, _, world = parallel.init_from_env("gloo")

class Videos:                      # stands in for CustomFrameDataset: only len() and the index order matter here
    def __len__(self): return 7
    def __getitem__(self, i): raise AssertionError("indices() must not touch the data")

train = list(FrameLoader(Videos(), shuffle=True, generator=torch.Generator().manual_seed(42), rank=rank, world_size=world).indices())
order = list(FrameLoader(Videos(), shuffle=True, generator=torch.Generator().manual_seed(42)).indices())
# training: every rank walks the same seed-42 order and takes every world-th video; the order is padded with its own head so
# that all ranks take the SAME number of steps (one gradient all-reduce per video: an odd video out would hang the others)
assert train == (order + order[:1])[rank::world] and len(train) == 4, (train, order)
mine = list(FrameLoader(Videos(), shuffle=False, rank=rank, world_size=world).indices())
assert mine == list(range(7))[rank::world]                 # validation: exact shards, no padding
# the only exchange of the frame / ensemble paths besides the gradient all-reduce: the sum of the confusion counts
counts = torch.tensor([rank + 1, 10 * (rank + 1), 0, 5], dtype=torch.int64)
parallel.allreduce_sum_(counts)
assert counts.tolist() == [3, 30, 0, 10], counts
got = [None] * world
dist.all_gather_object(got, mine)
assert sorted(sum(got, [])) == list(range(7))              # the shards partition the fold
dist.destroy_process_group()
print("ok", rank)
'''


def test_gloo_world2_frame_sharding_and_count_reduce(tmp_path):
    """N > 1 path of the frame / ensemble code on the host: videos are dealt round-robin over the ranks from one shared
    shuffled order, confusion counts are summed."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29741",
                   CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script), root], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_video_passes_partition():
    from multimodal_error_detection_b200.ensemble import video_passes
    assert video_passes([], 10) == []
    assert video_passes([3, 4, 5], 100) == [(0, 3)]
    assert video_passes([3, 4, 5], 7) == [(0, 2), (2, 3)]
    assert video_passes([30, 4, 5, 40, 1], 9) == [(0, 1), (1, 3), (3, 4), (4, 5)]       # over-long videos run alone
    lens = np.random.default_rng(0).integers(1, 50, size=200)
    passes = video_passes(lens, 120)
    assert passes[0][0] == 0 and passes[-1][1] == 200 and all(a[1] == b[0] for a, b in zip(passes, passes[1:]))
    assert all(lens[a:b].sum() <= 120 or b - a == 1 for a, b in passes)


def test_header_constants_match_the_binding():
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "b200med.h")).read()
    consts = dict(re.findall(r"#define (B200MED_TCN_\w+) (\d+)", header))
    from multimodal_error_detection_b200 import ops
    assert int(consts["B200MED_TCN_PACK_FLOATS"]) == ops.TCN_PACK_FLOATS == ref.PACK
    assert int(consts["B200MED_TCN_GRAD_FLOATS"]) == ops.TCN_GRAD_FLOATS == ref.GRAD


@pytest.mark.parametrize("tag", ["causal", "noncausal"])
def test_stage_node_matches_the_executed_reference(emulated, golden_dir, tag):
    """tests/golden/tecno_extra.* were recorded by EXECUTING the reference MultiStageModel (tests/golden/make_golden.py
    gen_tecno_extra): causal and non-causal, three videos one forward each.  The product model rebuilt from the same seed has
    the same weights (state digest); with the kernels replaced by their torch definitions it reproduces the reference's
    logits per video, its logits in ONE ragged pass over the three videos, and its gradients."""
    import json
    import os
    from hashing import state_digest
    meta = json.load(open(os.path.join(golden_dir, "tecno_extra.json")))
    gold = np.load(os.path.join(golden_dir, "tecno_extra.npz"))
    causal = tag == "causal"
    torch.manual_seed(meta["weight_seed"])
    m = models_TCN.MultiStageModel(meta["stages"], meta["layers"], meta["maps"], meta["dim"], meta["classes"], causal)
    assert state_digest(m.state_dict()) == meta[f"{tag}/state"]          # same construction order -> same seeded weights
    gen = torch.Generator().manual_seed(meta["input_seed"])
    videos = [torch.randn(1, n, meta["dim"], generator=gen) for n in meta["lengths"]]
    m.eval()
    with torch.no_grad():
        for i, x in enumerate(videos):
            out = m(x.permute(0, 2, 1))
            assert m.impl == "b200"
            assert np.abs(out.numpy() - gold[f"{tag}/logits{i}"]).max() <= 2e-5 * np.abs(gold[f"{tag}/logits{i}"]).max()
        ragged = m.forward_ragged(torch.cat([v[0] for v in videos]), meta["lengths"])       # [stages, C, sum T]
    want = np.concatenate([gold[f"{tag}/logits{i}"][:, 0] for i in range(len(videos))], axis=2)
    assert np.abs(ragged.numpy() - want).max() <= 2e-5 * np.abs(want).max()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m.train()
    x = videos[-1].clone().requires_grad_(True)
    y = m(x.permute(0, 2, 1))
    (y * torch.from_numpy(gold[f"{tag}/w"])).sum().backward()
    assert np.abs(x.grad.numpy() - gold[f"{tag}/dx"]).max() <= 1e-4 * np.abs(gold[f"{tag}/dx"]).max()
    names = [k for k, _ in m.named_parameters()]
    assert names == meta[f"{tag}/grad_names"]
    norms = np.asarray([float(p.grad.double().norm()) for _, p in m.named_parameters()])
    sums = np.asarray([float(p.grad.double().sum()) for _, p in m.named_parameters()])
    assert np.allclose(norms, gold[f"{tag}/grad_norms"], rtol=1e-4, atol=1e-6 * gold[f"{tag}/grad_norms"].max())
    assert np.allclose(sums, gold[f"{tag}/grad_sums"], rtol=1e-3, atol=1e-4 * gold[f"{tag}/grad_norms"].max())


def test_cabi_argument_errors_of_the_tcn_entry_points():
    """Error convention of the C ABI (include/b200med.h): bad arguments return B200MED_E_ARG (-1) with a message BEFORE any
    CUDA call is made -- checked here without a GPU; the Python binding maps -1 to ValueError."""
    import ctypes as C
    from multimodal_error_detection_b200 import _lib
    lib = _lib.load()
    null, one = C.c_void_p(0), C.c_void_p(256)            # `one`: a non-null, 16-byte aligned address that is never dereferenced
    rc = lib.b200med_tcn_layer_fwd(one, one, one, null, -1, 1, 1, null, null, 0.0, 0, null, 0, null)
    assert rc == -1 and b"bad shape" in lib.b200med_last_error()
    rc = lib.b200med_tcn_layer_fwd(one, one, one, null, 8, 0, 1, null, null, 0.0, 0, null, 0, null)
    assert rc == -1                                                                       # dilation 0
    rc = lib.b200med_tcn_layer_fwd(one, one, one, null, 8, 1, 1, null, null, 1.0, 0, null, 0, null)
    assert rc == -1 and b"dropout" in lib.b200med_last_error()
    rc = lib.b200med_tcn_layer_fwd(null, one, one, null, 8, 1, 1, null, null, 0.0, 0, null, 0, null)
    assert rc == -1 and b"null pointer" in lib.b200med_last_error()
    rc = lib.b200med_tcn_layer_fwd(C.c_void_p(260), one, one, null, 8, 1, 1, null, null, 0.0, 0, null, 0, null)
    assert rc == -1 and b"aligned" in lib.b200med_last_error()
    assert lib.b200med_tcn_layer_fwd(one, one, one, null, 0, 1, 1, null, null, 0.0, 0, null, 0, null) == 0   # T = 0: nothing to do
    rc = lib.b200med_tcn_out_fwd(one, one, one, one, 8, 9, null)
    assert rc == -1 and b"classes" in lib.b200med_last_error()
    rc = lib.b200med_tcn_layer_bwd_hidden(one, one, one, one, one, one, 0, 8, 1, 1, null, null, 0.0, 0, null, 0, null)
    assert rc == -1                                                                       # zero slots
    assert lib.b200med_tcn_slots(600) == 75 and lib.b200med_tcn_slots(641) == 41 and lib.b200med_tcn_slots(5000) == 80
    with pytest.raises(ValueError, match="bad shape"):
        _lib.call("b200med_tcn_reduce_grads", one, 0, 1, one, null)
