"""Per-kernel roofline table of one ImageGPT C5 transformer block (P = 64*1024 pixels, C = 512, 8 heads x 64):
algorithmic flops / HBM bytes -> ideal time = max(flops / tensor peak, bytes / HBM peak), against the measured
microbenchmarks in profiles/ (r01_gemm_microbench_final.txt + stored-derivative run, r01_ln_microbench.txt,
attention ncu captures).  Peaks: MEASURED_PEAKS.json (sustained bf16, HBM copy)."""
TF, GBS = 1447.6e12, 6574.8e9
P, C, H, D, S, N = 65536, 512, 8, 64, 1024, 64
bf, f32 = 2, 4
rows = []

def gemm(name, m, n, k, bytes_, measured_us):
    fl = 2.0 * m * n * k
    rows.append((name, fl, bytes_, measured_us))

X = P * C  # elements of a [P, 512] tensor
gemm("qkv fwd (bias)", P, 3 * C, C, X * bf + 3 * X * bf, 103.4)
gemm("proj fwd (bias, res -> fp32)", P, C, C, X * bf + X * f32 + X * f32, 68.6)
gemm("fc1 fwd (bias, GELU, GELU')", P, 4 * C, C, X * bf + 2 * 4 * X * bf, 199.8)
gemm("fc2 fwd (bias, 2 res -> fp32)", P, C, 4 * C, 4 * X * bf + 2 * X * f32 + X * f32, 144.4)
gemm("fc2 dgrad (x stored GELU')", P, 4 * C, C, X * bf + 4 * X * bf + 4 * X * bf, 163.9)
gemm("fc1 dgrad", P, C, 4 * C, 4 * X * bf + X * bf, 99.3)
gemm("qkv dgrad", P, C, 3 * C, 3 * X * bf + X * bf, 80.6)
gemm("proj dgrad", P, C, C, 2 * X * bf, 43.0)
gemm("fc1 wgrad (split-K 4)", 4 * C, C, P, 4 * X * bf + X * bf, 113.7)
gemm("fc2 wgrad (split-K 4)", C, 4 * C, P, 4 * X * bf + X * bf, 113.7)
gemm("qkv wgrad (split-K 8)", 3 * C, C, P, 3 * X * bf + X * bf, 114.7)
gemm("proj wgrad (split-K 32)", C, C, P, 2 * X * bf, 48.1)
tiles = N * H * (S // 128) * (S // 128 + 1) // 2 * 128 * 128  # (q, k) pairs at tile granularity
rows.append(("attention fwd (exp2 floor 67 us)", 4.0 * D * tiles, 4 * X * bf, 191.6))
rows.append(("attention bwd", 10.0 * D * tiles, 5 * X * bf + 3 * X * bf + 2 * X * f32, 517.0))
rows.append(("LayerNorm fwd (x2)", 0, X * (f32 + bf), 38.9))
rows.append(("LayerNorm-2 bwd (1 residual grad)", 0, X * (bf + f32 + f32 + f32 + bf), 106.5))
rows.append(("LayerNorm-1 bwd (2 residual grads)", 0, X * (bf + f32 + f32 + f32 + f32 + bf), 119.8))
rows.append(("bias-grad column sums (dqkv, dh)", 0, 3 * X * bf + 4 * X * bf, 2 * 38.6))

print("| kernel | GFLOP | HBM MB | ideal us (bound) | measured us | fraction |")
print("|---|---|---|---|---|---|")
ti = tm = 0.0
for name, fl, by, us in rows:
    t_f, t_b = fl / TF * 1e6, by / GBS * 1e6
    ideal = max(t_f, t_b)
    mult = 2 if "(x2)" in name else 1
    ti += ideal * mult
    tm += us * mult
    print(f"| {name} | {fl / 1e9:.1f} | {by / 1e6:.0f} | {ideal:.1f} ({'tensor' if t_f >= t_b else 'HBM'}) | {us:.1f} | {ideal / us:.2f} |")
print(f"| **block total** | | | {ti:.0f} | {tm:.0f} | {ti / tm:.2f} |")
print(f"\n24 blocks: ideal {24 * ti / 1e3:.1f} ms, measured kernels {24 * tm / 1e3:.1f} ms (step 69.4 ms incl. Adam, clip, fills, input/output layers, host gaps)")
