"""TMA-row cost model of conv_umma_kernel vs measured per-layer times (gpurun_out/layers_*.txt)."""
import re, sys
f = sys.argv[1]
CPR = float(sys.argv[2]) if len(sys.argv) > 2 else 2.2   # cycles per TMA row
GHZ, SMS = 1.9, 148
rows = [l.split(None, 2) for l in open(f)]
seen = {}
for ms, gf, tag in rows:
    tag = tag.strip()
    if not tag.startswith("conv"): continue
    seen.setdefault(tag, []).append(float(ms))
tot_meas = tot_mma = tot_row = tot_model = 0
print(f"{'layer':44s} {'meas':>6s} {'mma':>6s} {'rows':>6s} {'A%':>4s}  opts: halo  mc4   both")
for tag, mss in seen.items():
    m = re.match(r"conv n(\d+) (\d+)x(\d+) cin(\d+) cout(\d+) k(\d+)x(\d+) s(\d+)", tag)
    n, h, w, cin, cout, R, S, st = map(int, m.groups())
    # output dims are not in the tag: infer pad from kernel (inception: 'same' for odd k at s1 except stem/no-pad layers)
    nopad = (h, cin, cout) in [(149, 32, 32), (73, 80, 192)] or st == 2
    ho = (h - R) // st + 1 if nopad else h
    wo = (w - S) // st + 1 if nopad else w
    M = n * ho * wo
    nt = max(t for t in range(16, 257, 16) if cout % t == 0)
    tiles_m = -(-M // 128); ntl = cout // nt
    nch = -(-cin // 64); kb = R * S * nch
    mma = tiles_m * ntl * kb * 4 * 2 * (nt / 2)            # cycles (2 passes, 4 ksteps of nt/2 cycles), upper bound (full chunks)
    mma = tiles_m * 128 * cout * R * S * cin * 2 / 4096      # exact MAC count, 2 passes
    a_rows = tiles_m * ntl * kb * 128
    b_rows = tiles_m * ntl * kb * 2 * nt
    a_halo = tiles_m * ntl * nch * R * (128 + S - 1) if st == 1 else a_rows
    def t(ar, br): return max(mma, (ar + br) * CPR) / SMS / GHZ / 1e3
    meas = sum(mss) / len(mss) * 1000
    cnt = len(mss) // 8 if len(mss) >= 8 else 1
    model = t(a_rows, b_rows)
    print(f"{tag[5:]:44s} {meas:6.0f} {mma/SMS/GHZ/1e3:6.0f} {model:6.0f} {100*a_rows/(a_rows+b_rows):4.0f}       {t(a_halo,b_rows):5.0f} {t(a_rows,b_rows/4):5.0f} {t(a_halo,b_rows/4):5.0f}  x{cnt}")
    tot_meas += meas * cnt; tot_mma += mma/SMS/GHZ/1e3 * cnt; tot_row += model * cnt
    tot_model += 0
print("totals (us per microbatch): measured %.0f  mma-ideal %.0f  row-model %.0f" % (tot_meas, tot_mma, tot_row))
