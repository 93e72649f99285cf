"""Per-kernel roofline table of one ImageGPT C5 transformer block (P = 64*1024 pixels, C = 512, 8 heads x 64):
algorithmic flops / HBM bytes read / written -> ideal time = max(flops / tensor peak, (R + W) / HBM copy peak,
W / HBM write peak), against the measured microbenchmarks of round 2 (profiles/r02_gemm_microbench.txt and the later
A/B runs, r02_attn_microbench.txt, r02_ln_gated_microbench.txt).  Peaks: MEASURED_PEAKS.json (sustained bf16, HBM copy);
the write-only peak is this repo's own measurement (tools/micro/write_bw.py: a 537 MB memset runs at 3.72 TB/s on the same
boxes whose copy runs at 6.24 TB/s of read + write) — a kernel that mostly WRITES is bound by that, not by the copy peak."""
TF, GBS, GBS_W = 1447.6e12, 6574.8e9, 3720e9
P, C, H, D, S, N = 65536, 512, 8, 64, 1024, 64
bf, f32 = 2, 4
rows = []
X = P * C  # elements of a [P, 512] tensor
W1 = C * C * bf


def gemm(name, m, n, k, rd, wr, measured_us):
    rows.append((name, 2.0 * m * n * k, rd, wr, measured_us))


gemm("qkv fwd (bias)", P, 3 * C, C, X * bf + 3 * W1, 3 * X * bf, 93.2)
gemm("proj fwd (bias, res -> fp32)", P, C, C, X * bf + X * f32 + W1, X * f32, 66.6)
gemm("fc1 fwd (bias, GELU, GELU')", P, 4 * C, C, X * bf + 4 * W1, 2 * 4 * X * bf, 191.5)
gemm("fc2 fwd (bias, 2 res -> fp32)", P, C, 4 * C, 4 * X * bf + 2 * X * f32 + 4 * W1, X * f32, 141.3)
gemm("fc2 dgrad (x stored GELU')", P, 4 * C, C, X * bf + 4 * X * bf + 4 * W1, 4 * X * bf, 152.4)
gemm("fc1 dgrad", P, C, 4 * C, 4 * X * bf + 4 * W1, X * bf, 91.2)
gemm("qkv dgrad", P, C, 3 * C, 3 * X * bf + 3 * W1, X * bf, 75.8)
gemm("proj dgrad", P, C, C, X * bf + W1, X * bf, 41.0)
gemm("fc1 wgrad (split-K 4)", 4 * C, C, P, 4 * X * bf + X * bf, 4 * W1 * 2, 113.5)
gemm("fc2 wgrad (split-K 4)", C, 4 * C, P, 4 * X * bf + X * bf, 4 * W1 * 2, 113.5)
gemm("qkv wgrad (split-K 8)", 3 * C, C, P, 3 * X * bf + X * bf, 3 * W1 * 2, 113.7)
gemm("proj wgrad (split-K 32)", C, C, P, 2 * X * bf, W1 * 2, 47.1)
tiles = N * H * (S // 128) * (S // 128 + 1) // 2 * 128 * 128  # (q, k) pairs at tile granularity
rows.append(("attention fwd (exp2 floor 67 us)", 4.0 * D * tiles, 3 * X * bf, X * bf, 206.8))
rows.append(("attention bwd (+ delta, dQ zero / convert)", 10.0 * D * tiles, 5 * X * bf + X * f32, 3 * X * bf + 2 * X * f32, 492.6))
rows.append(("LayerNorm fwd (x2)", 0, X * f32, X * bf, 38.9))
rows.append(("LayerNorm-2 bwd (1 residual grad)", 0, X * (bf + f32 + f32), X * (f32 + bf), 104.5))
rows.append(("LayerNorm-1 bwd (2 residual grads)", 0, X * (bf + f32 + f32 + f32), X * (f32 + bf), 118.8))
# the bias gradients of the qkv and fc1 layers ride on their wgrad launches since the end of round 2 (before: two
# column-sum passes, 470 MB read, 77 us per block)

print("| kernel | GFLOP | read MB | written MB | ideal us (bound) | measured us | fraction |")
print("|---|---|---|---|---|---|---|")
ti = tm = 0.0
for name, fl, rd, wr, us in rows:
    t_f, t_b, t_w = fl / TF * 1e6, (rd + wr) / GBS * 1e6, wr / GBS_W * 1e6
    ideal = max(t_f, t_b, t_w)
    bound = "tensor" if ideal == t_f else ("HBM r+w" if ideal == t_b else "HBM write")
    mult = 2 if "(x2)" in name else 1
    ti += ideal * mult
    tm += us * mult
    print(f"| {name} | {fl / 1e9:.1f} | {rd / 1e6:.0f} | {wr / 1e6:.0f} | {ideal:.1f} ({bound}) | {us:.1f} | {ideal / us:.2f} |")
print(f"| **block total** | | | | {ti:.0f} | {tm:.0f} | {ti / tm:.2f} |")
print(f"\n24 blocks: ideal {24 * ti / 1e3:.1f} ms, measured kernels {24 * tm / 1e3:.1f} ms (step 60.5 ms incl. Adam, clip, input / output layers, host gaps)")
