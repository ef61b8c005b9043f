"""Comparator script (test infrastructure, lives under tests/ because it executes the oracle; NOT the product, not a bench.py
arm): the reference algorithm as PyTorch eager on the SAME B200 — the
oracle port (oracle/onet_oracle.py: the ATen ops the reference module dispatches to; on a CUDA device these are
cuDNN / cuBLAS kernels) running the training step of Train_Onet_on_simclutter_20250407.py:209-218 on the benchmark
batch, in FP32 (TF32 tensor cores allowed) and under bf16 autocast with channels_last activations.  SURVEY.md §8d asks
for it as "the honest GPU comparator" next to the CPU baseline.

    python tests/compare_torch_gpu.py [--batch 64] [--steps 5]   ->  one JSON line per precision
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--device", default="cuda")
    args = ap.parse_args()
    from oracle import onet_oracle as orc
    from onet_b200.data import k_clutter_frames
    dev = torch.device(args.device)
    x = k_clutter_frames(args.batch, 1, args.size, args.size, seed=7).to(dev)
    for prec in ("fp32_tf32", "bf16_autocast"):
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.benchmark = True
        st = {k: v.to(dev) for k, v in orc.init_state(1, seed=1981).items()}
        leaves = [k for k, v in st.items() if v.dtype.is_floating_point and "running" not in k]
        params = [st[k].requires_grad_(True) for k in leaves]
        opt = torch.optim.Adam(params, lr=5e-6, fused=dev.type == "cuda")
        xin = x.contiguous(memory_format=torch.channels_last) if prec == "bf16_autocast" else x

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast(dev.type, dtype=torch.bfloat16, enabled=prec == "bf16_autocast"):
                Lt, Vt, Ld, Vd, S = orc.onet_forward(st, xin, training=True)
            loss = orc.compute_loss(Lt.float(), S[:, 0:1].float(), Ld.float(), S[:, 1:2].float())
            loss.backward()
            opt.step()
            return loss

        try:
            for _ in range(args.warmup):
                step()
            if dev.type == "cuda":
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    loss = step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.steps
                mem = torch.cuda.max_memory_allocated() / 2 ** 30
            else:
                import time
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    loss = step()
                ms = (time.perf_counter() - t0) * 1e3 / args.steps
                mem = None
            print(json.dumps(dict(comparator="pytorch eager (oracle port on ATen/cuDNN)", precision=prec, batch=args.batch,
                                  image=[1, args.size, args.size], images_per_s=args.batch / (ms * 1e-3), ms_per_step=ms,
                                  loss=float(loss), peak_mem_gib=mem, torch=torch.__version__)), flush=True)
        except Exception as e:      # e.g. out of memory at this batch in fp32
            print(json.dumps(dict(comparator="pytorch eager", precision=prec, batch=args.batch, error=repr(e)[:300])), flush=True)
        del st, params, opt
        if dev.type == "cuda":
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
