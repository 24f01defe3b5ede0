"""Frame path (BASELINE config 0, train_frame: TeCNo over per-frame features) -- A/B of the fused TeCNo kernels
(csrc/tcn.cu) against the same model on the stock torch CUDA convolution layers, one video per step like the
reference's DataLoader(batch_size=1), plus ragged-batched inference.  CUDA-event timing, warm-up first.

    python scripts/bench_frame.py [--frames 600] [--videos 64] [--steps 30]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_error_detection_b200 import _lib  # noqa: E402
from multimodal_error_detection_b200.modeling import modeling_utils as mu  # noqa: E402


def timed(fn, steps, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def torch_layers_forward(model, x):
    """The A/B baseline of this script: the SAME parameters run through stock torch CUDA convolutions (what the reference's
    nn.Module forward does, models_TCN.py:76-137).  It lives here, not in the package: the product has one implementation."""
    import torch.nn.functional as F

    def stage(s, h):
        h = F.conv1d(h, s.conv_1x1.weight, s.conv_1x1.bias)
        for l in s.layers:
            d = l.dilation
            y = F.relu(F.conv1d(h, l.conv_dilated.weight, l.conv_dilated.bias, padding=2 * d if l.causal_conv else d, dilation=d))
            if l.causal_conv:
                y = y[:, :, :-(2 * d)]
            h = h + F.dropout(F.conv1d(y, l.conv_1x1.weight, l.conv_1x1.bias), 0.5, model.training)
        return F.conv1d(h, s.conv_out_classes.weight, s.conv_out_classes.bias)

    out = stage(model.stage1, x)
    outs = [out]
    for s in model.stages:
        out = stage(s, F.softmax(out, dim=1))
        outs.append(out)
    return torch.stack(outs, dim=0)


def measure(frames: int = 600, videos: int = 64, steps: int = 30) -> dict:
    args = argparse.Namespace(frames=frames, videos=videos, steps=steps)
    dev = torch.device("cuda", torch.cuda.current_device())
    kw = dict(dataset_type="frame", error_type="global", pos_weight=False, n_epochs=2, batch_size=1, lr=3e-4,
              lr_scheduler=True, weight_decay=1e-4, num_layers=3, hidden_size=128, video_dims=32, data_type="multimodal",
              delete_ND=True, return_train_preds=False, siamese=False, model_name="TeCNo", mstcn_stages=2, mstcn_layers=8,
              mstcn_f_maps=64, mstcn_f_dim=58, out_features=2, mstcn_causal_conv=True)
    fe, model, crit, opt, sched = mu.define_model_objects(kw, {"multimodal": 58, "video": 32, "kinematics": 26}, dev, (0.4, 0.6))
    T = args.frames
    g = torch.Generator().manual_seed(42)
    images = torch.randn(1, T, 2048, generator=g).clamp_min(0).to(dev)
    kin = torch.randn(1, T, 26, generator=g).to(dev)
    y = (torch.rand(1, T, generator=g) > 0.5).float().to(dev)
    res = {"frames_per_video": T, "config": "TeCNo 2 stages x 8 layers x 64 maps, F=58 (FE 2048->32 + 26 kinematics), fp32"}

    fwd = {"fn": model}

    def train_step():
        inputs = mu.define_inputs(images, kin, fe, kw, dev)
        out = fwd["fn"](inputs)
        loss, _ = mu.compute_loss(out, y, crit, "frame")
        opt.zero_grad()
        loss.backward()
        opt.step()

    def head_only():
        out = fwd["fn"](feats)
        (out.sum()).backward()

    def infer():
        with torch.no_grad():
            fwd["fn"](mu.define_inputs(images, kin, fe, kw, dev))

    feats = torch.randn(1, T, 58, generator=g).to(dev).permute(0, 2, 1)
    for name, fused in (("b200", True), ("torch_layers", False)):
        fwd["fn"] = model if fused else (lambda x: torch_layers_forward(model, x))
        model.train(); fe.train()
        n0 = _lib.launch_count()
        ms = timed(train_step, args.steps)
        launches = (_lib.launch_count() - n0) / (args.steps + 5)
        ms_head = timed(head_only, args.steps)
        model.eval(); fe.eval()
        ms_inf = timed(infer, args.steps)
        res[name] = {"train_ms_per_video": ms, "train_frames_per_s": T / ms * 1e3, "head_fwd_bwd_ms": ms_head,
                     "infer_ms_per_video": ms_inf, "infer_frames_per_s": T / ms_inf * 1e3, "own_launches_per_step": launches,
                     "impl": "b200" if fused else "torch layers (script-local baseline)"}
    # the same train step replayed from a CUDA graph (engine.FrameTrainStep; what train_single_epoch does with cuda_graph=True)
    from multimodal_error_detection_b200.engine import FrameTrainStep
    model.train(); fe.train()
    e7 = torch.zeros(1, T, 7, dtype=torch.int32, device=dev)
    e7[0, :, 6] = y[0].to(torch.int32)
    step = FrameTrainStep(images, kin, e7, fe, model, crit, opt, kw).capture()
    ms = timed(step.run, args.steps)
    res["b200_graph"] = {"train_ms_per_video": ms, "train_frames_per_s": T / ms * 1e3, "own_launches_per_step": step.launches_per_step}
    # the bf16 mode of the same step: FeatureExtractor on the tcgen05 GEMMs (2e-2 bar), TeCNo training on its fp32 kernels
    try:
        kw16 = dict(kw, precision="bf16")
        fe16, model16, crit16, opt16, _ = mu.define_model_objects(kw16, {"multimodal": 58, "video": 32, "kinematics": 26}, dev, (0.4, 0.6))
        model16.train(); fe16.train()
        step16 = FrameTrainStep(images, kin, e7, fe16, model16, crit16, opt16, kw16).capture()
        ms16 = timed(step16.run, args.steps)
        res["b200_graph_bf16_fe"] = {"train_ms_per_video": ms16, "train_frames_per_s": T / ms16 * 1e3,
                                     "own_launches_per_step": step16.launches_per_step, "loss": float(step16.loss.item())}
        del step16
    except Exception as e:      # noqa: BLE001 -- an auxiliary line must not take the others down
        res["b200_graph_bf16_fe"] = {"error": f"{type(e).__name__}: {e}"}
    # ragged-batched inference of the head: V videos in one pass vs one pass per video
    model.eval()
    lengths = torch.randint(300, 901, (args.videos,), generator=g).tolist()
    frames = torch.randn(sum(lengths), 58, generator=g).to(dev)

    def ragged():
        model.forward_ragged(frames, lengths)

    def per_video():
        s = 0
        with torch.no_grad():
            for n in lengths:
                model(frames[s:s + n].unsqueeze(0).permute(0, 2, 1))
                s += n

    ms_r, ms_p = timed(ragged, 10, 2), timed(per_video, 3, 1)
    res["head_inference"] = {"videos": args.videos, "frames": sum(lengths), "ragged_ms": ms_r, "per_video_ms": ms_p,
                             "ragged_frames_per_s": sum(lengths) / ms_r * 1e3, "per_video_frames_per_s": sum(lengths) / ms_p * 1e3}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=600)
    ap.add_argument("--videos", type=int, default=64)
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    print(json.dumps(measure(a.frames, a.videos, a.steps)))


if __name__ == "__main__":
    main()
