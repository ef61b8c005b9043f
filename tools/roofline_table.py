"""Per-launch roofline table of one training step from bench.py's per-call dump (ONET_BENCH_DETAIL=...tsv):

    python tools/roofline_table.py profiles/r1_step_detail_v5.tsv > profiles/r1_per_layer_roofline_v5.md

For every C-ABI call of the step: algorithmic FLOPs F and algorithmic HBM bytes B (what the operation must read and write once,
bf16 activations), the roofline time max(F / peak_tensor, B / peak_hbm) with the MEASURED peaks of
MEASURED_PEAKS.json (sustained bf16 matmul, copy bandwidth), and the fraction roofline time / measured time.  The last rows
give the step totals: the sum of the roofline times is the time an ideal implementation of the SAME kernel decomposition would
need; fusing passes away (fewer bytes) is a different roofline and is discussed in DESIGN.md §8."""
import ast
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def work(name, a, e=2.0):
    """(description, flops, bytes) of one call from its integer arguments (pointers are not in the dump); e = bytes per
    activation element (2 = bf16 mode, 4 = fp32 / tf32 modes)."""
    if name == "onet_conv3x3_fwd":
        ldi, _, n, h, w, ci, co = a[:7]
        return f"conv3x3 {ci}->{co} @{h}x{w}", 2.0 * 9 * n * h * w * ci * co, n * h * w * (ci + co) * e + 9 * ci * co * e
    if name == "onet_conv3x3_dgrad_bnred":
        ldi, _, n, h, w, ci, co = a[:7]
        return f"conv3x3 dgrad+bnred {ci}->{co} @{h}x{w}", 2.0 * 9 * n * h * w * ci * co, n * h * w * (ci + 2 * co) * e + 9 * ci * co * e
    if name == "onet_conv3x3_wgrad":
        _, _, _, _, n, h, w, ci, co = a[:9]
        return f"conv3x3 wgrad {ci}->{co} @{h}x{w}", 2.0 * 9 * n * h * w * ci * co, n * h * w * (ci + co) * e + 9 * ci * co * 4
    if name == "onet_convT2x2_fwd":
        _, _, n, h, w, ci, co = a[:7]
        return f"up-conv {ci}->{co} @{h}x{w}", 2.0 * 4 * n * h * w * ci * co, n * h * w * (ci + 4 * co) * e
    if name == "onet_convT2x2_dgrad":
        _, _, n, h, w, ci, co = a[:7]
        return f"up-conv dgrad {ci}->{co} @{h}x{w}", 2.0 * 4 * n * h * w * ci * co, n * h * w * (ci + 4 * co) * e
    if name == "onet_convT2x2_wgrad":
        _, _, _, _, n, h, w, ci, co = a[:9]
        return f"up-conv wgrad {ci}->{co} @{h}x{w}", 2.0 * 4 * n * h * w * ci * co, n * h * w * (ci + 4 * co) * e
    if name == "onet_first_conv_stats":
        n, h, w, ci = a[:4]
        return f"first conv {ci}->64 @{h}x{w}: statistics (recomputed)", 0.0, n * h * w * ci * e
    if name == "onet_first_conv_bn_relu":
        n, h, w, ci = a[:4]
        return f"first conv {ci}->64 @{h}x{w}: conv + BN + ReLU", 0.0, n * h * w * (ci + 64) * e
    if name == "onet_first_conv_bwd":
        n, h, w, ci = a[:4]
        # algorithmic minimum: one pass over g and x (what the in_chns = 1 closed-form path does; the two-pass form reads them twice)
        return f"first conv {ci}->64 @{h}x{w}: BN backward + wgrad (recomputed)", 0.0, n * h * w * (ci + 64) * e
    if name == "onet_bn_relu_apply":
        n, h, w, c, _, ldo = a[:6]
        pooled = ldo == 2 * c
        return f"BN+ReLU{'+pool' if pooled else ''} C={c} @{h}x{w}", 0.0, n * h * w * c * (2 * e + (e / 4 if pooled else 0))
    if name == "onet_bn_relu_bwd":
        n, h, w, c, _, ld1, _, ld2 = a[:8]
        pooled, g2 = ld1 == 2 * c, ld2 != 0
        reads = e + e + (e if g2 else 0) + (e / 4 if pooled else 0)
        return f"BN+ReLU{'+pool' if pooled else ''}{'+2nd grad' if g2 else ''} backward C={c} @{h}x{w}", 0.0, n * h * w * c * (2 * reads + e)
    if name == "onet_bn_relu_bwd_apply":
        n, h, w, c = a[:4]
        return f"BN+ReLU backward (apply only) C={c} @{h}x{w}", 0.0, n * h * w * c * 3 * e
    if name == "onet_head_fwd_bn":
        b, h, w = a[4:7]
        return f"head + JSD loss (+ last BN + ReLU) @{h}x{w}", 0.0, b * h * w * (4 * 64 * e + 24)
    if name == "onet_head_bwd_scalars":
        b, h, w = a[:3]
        return f"head backward, per-pixel part @{h}x{w}", 0.0, b * h * w * 32.0
    if name == "onet_bn_relu_bwd_head":
        n, h, w, c = a[:4]
        return f"last BN+ReLU backward fused with head backward C={c} @{h}x{w}", 0.0, n * h * w * c * 6 * e
    if name == "onet_head_fwd":
        b, h, w = a[4:7]
        return f"head + JSD loss @{h}x{w}", 0.0, b * h * w * (4 * 64 * e + 24)
    if name == "onet_head_bwd":
        b, h, w = a[4:7]
        return f"head backward @{h}x{w}", 0.0, b * h * w * (4 * 64 * e + 16 + 4 * 64 * e)
    if name == "onet_adam_step_dev":
        return "Adam", 0.0, a[0] * 28.0
    if name == "onet_pack_all_weights":
        return "pack weights (fp32 -> bf16 operands)", 0.0, 31036416 * 8.0
    if name == "onet_prep_input":
        b, c, h, w = a[:4]
        return "twin input (X, 1-X)", 0.0, b * c * h * w * (4 + 2 * e)
    return name.replace("onet_", ""), 0.0, 0.0


def main():
    path = sys.argv[1]
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(ppath):
        pk = json.load(open(ppath))
    else:          # fallback figures of the profiling recipe (B200_PROFILING.md), as in bench.py
        pk = dict(bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, hbm_gbs=6650.0)
    tf, bw = pk["bf16_tflops_sustained"], pk["hbm_gbs"]
    print(f"# Per-launch roofline of one training step ({os.path.basename(path)})\n")
    print(f"Peaks ({'MEASURED_PEAKS.json' if os.path.isfile(ppath) else 'fallback of the profiling recipe'}): bf16 {tf} TFLOP/s sustained ({pk['bf16_tflops']} burst), HBM copy {bw} GB/s.  "
          "B = 64 frames per GPU (128 twin images), 1x256x256, bf16.  Times: CUDA events around every call, serial schedule "
          "(the profile pass of bench.py).\n")
    print("| # | call | kernel | ms | TFLOP/s | GB/s | bound | roofline ms | fraction |")
    print("|---|---|---|---|---|---|---|---|---|")
    tot_ms = tot_roof = 0.0
    by = {}
    for i, ln in enumerate(open(path)):
        f = ln.rstrip("\n").split("\t")
        name, kern, ms, ints = f[0], f[1], float(f[3]), ast.literal_eval(f[5])
        desc, fl, by_ = work(name, ints)
        t_t, t_b = fl / (tf * 1e12) * 1e3, by_ / (bw * 1e9) * 1e3
        roof = max(t_t, t_b)
        bound = "tensor" if t_t >= t_b else "hbm"
        if roof == 0.0:
            bound = "-"
        tot_ms += ms
        tot_roof += roof
        d = by.setdefault(bound, [0.0, 0.0])
        d[0] += ms
        d[1] += roof
        print(f"| {i} | {desc} | `{kern}` | {ms:.3f} | {fl / ms / 1e9 if fl else 0:.0f} | {by_ / ms / 1e6 if by_ else 0:.0f} | {bound} | "
              f"{roof:.3f} | {roof / ms if ms else 0:.2f} |")
    print(f"\n**Step (serial sum of the calls): {tot_ms:.2f} ms measured, {tot_roof:.2f} ms roofline = {tot_roof / tot_ms:.2f}.**  ")
    for b, (m, r) in sorted(by.items()):
        if b != "-":
            print(f"{b}-bound calls: {m:.2f} ms measured, {r:.2f} ms roofline ({r / m:.2f}).  ")
    print("\nThe replayed step (graph, two-stream backward) is shorter than the serial sum: see the bench line of the same round.")


if __name__ == "__main__":
    main()
