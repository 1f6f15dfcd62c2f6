"""Per-layer bounds of the conv launches of one step, from a tuned per-op table (tools/profile_ops.py output):
tensor pipe, SHARED-MEMORY bandwidth (every byte a work item moves through the SM's 128 B/clk shared-memory port: TMA
fills, the operand reads of every tcgen05.mma - 128 x 16 + N x 16 bf16 per instruction -, the epilogue's staging tile
and its TMA store / residual load), L2->SM operand ingest and HBM (unique input + output bytes), next to the measured time.

    python tools/layer_rooflines.py profiles/r2_ops_b64_tuned.txt > profiles/r2_layer_rooflines.txt

Peaks: tensor 2250 TFLOP/s nominal dense bf16 (1381 measured sustained cuBLAS, MEASURED_PEAKS.json); L2->SM 8.8 TB/s =
the plateau of l1tex__m_xbar2l1tex_read_bytes / duration and lts__t_bytes / duration over the conv launches
(profiles/r1_xr_ncu_*); HBM 6.44 TB/s measured copy bandwidth."""
import math
import re
import sys

B = 32
T_NOM, T_SUS, L2SM, HBM = 2250e12, 1381e12, 8.8e12, 6.44e12
SMEM = 148 * 128 * 1.9e9   # bytes/s through the shared-memory ports of 148 SMs at 1.9 GHz (128 B/clk/SM)
pat = re.compile(r"^(\S+)\s+([\d.]+) us\s+([\d.]+) TF/s\s+[\d.]+%\s+k(\d) s(\d) cin\s*(\d+) cout\s*(\d+) in\s*(\d+) mt(-?\d+) st(\d+) bn(\d+) bk(\d+) tile(\d+)x(\d+)")
rows = []
for line in open(sys.argv[1]):
    hb = re.match(r"^B=(\d+)", line)
    if hb:
        B = int(hb.group(1))
    m = pat.match(line)
    if not m:
        continue
    label, us = m.group(1), float(m.group(2))
    k, s, cin, cout, res, mt, st, bn, bk, tw, th = (int(x) for x in m.groups()[3:])
    if res > 5000:   # survivor-patch launches: work depends on the survivors
        continue
    Ho = res // s
    taps, cblks, npix = k * k, cin // bk, tw * th
    tiles = math.ceil(Ho / tw) * math.ceil(Ho / th) * B
    up = label.endswith(".up")
    flops = 2.0 * cin * cout * Ho * Ho * B * (1 if up else taps)
    es_out = 4 if ("preds" in label or "flame_out" in label) else 2
    hbm = B * res * res * cin * 2 + B * Ho * Ho * cout * es_out * (1 if not up else 1)
    if mt < 0:  # operand-swapped kernel: M = 128 weight rows of a group, N = pixel tile
        G = math.ceil(cout / 128)
        gw = cout // G
        items = tiles * G
        xr = st >= 100
        pair = mt <= -20   # cta_group::2: each CTA of a pair stages half a pixel tile and supplies N/2 operand rows per MMA
        if xr and pair:
            ingest_item = cblks * (3 * (th // 2 + 2) * tw * bk * 2 + 9 * gw * bk * 2)
        elif xr:
            ingest_item = cblks * (3 * (th + 2) * tw * bk * 2 + 9 * gw * bk * 2)
        else:
            ingest_item = taps * cblks * (npix * bk * 2 + gw * bk * 2)
        mma_item = taps * cblks * (bk // 16) * 2.0 * 128 * npix * 16
        n_mma = taps * cblks * (bk // 16)
        smem_item = ingest_item + n_mma * (128 + (npix // 2 if pair else npix)) * 32 + 2 * npix * gw * es_out   # fills + MMA operand reads + staging write / store read
        kind = "pair+XR" if pair else "swap+XR" if xr else "swap"
    else:       # pixels on M (mt tiles of 128), Cout tile on N
        n_tiles = math.ceil(cout / bn)
        items = math.ceil(B * Ho * Ho / (128 * mt)) * n_tiles
        ingest_item = taps * cblks * (mt * 128 * bk * 2 + bn * bk * 2)
        mma_item = taps * cblks * (bk // 16) * mt * 2.0 * 128 * bn * 16
        smem_item = ingest_item + taps * cblks * (bk // 16) * mt * (128 + bn) * 32   # (epilogue goes from registers to global)
        kind = f"normal mt{mt}"
    if ".cv2" in label:
        hbm += B * Ho * Ho * cout * 2
        ingest_item += (npix if mt < 0 else 128 * mt) * (gw if mt < 0 else bn) * 2  # residual tile
        if mt < 0:
            smem_item += 2 * npix * gw * 2   # residual tile: TMA fill + read-modify-write pass
    t_mma = items * mma_item / T_NOM
    t_ing = items * ingest_item / L2SM
    t_hbm = hbm / HBM
    t_smem = items * smem_item / SMEM
    bound = max((t_mma, "tensor"), (t_smem, "smem"), (t_hbm, "HBM"))
    waves = items / 148.0
    rows.append((label, kind, f"{tw}x{th}", us, flops / us / 1e6, t_mma * 1e6, t_smem * 1e6, t_ing * 1e6, t_hbm * 1e6, bound[1], bound[0] * 1e6 / us, waves))
print(f"# per-layer bounds (us) of the dense conv launches, batch {B}, 640x640; 'of bound' = max(tensor, smem, HBM) / measured.")
print("# tensor = executed MMA work (incl. padded rows / overhanging tiles) at the nominal 2250 TFLOP/s; smem = bytes through the")
print("# shared-memory port at 128 B/clk/SM, 1.9 GHz; L2->SM (8.8 TB/s plateau of round 1) is listed for reference only.")
print(f"{'layer':28s} {'kernel':10s} {'tile':>6s} {'meas us':>8s} {'TF/s':>7s} {'tensor':>7s} {'smem':>7s} {'L2->SM':>7s} {'HBM':>6s}  {'bound':7s} {'of bound':>8s} {'waves':>6s}")
tot = [0.0] * 5
for r in rows:
    print(f"{r[0]:28s} {r[1]:10s} {r[2]:>6s} {r[3]:8.1f} {r[4]:7.0f} {r[5]:7.1f} {r[6]:7.1f} {r[7]:7.1f} {r[8]:6.1f}  {r[9]:7s} {r[10]:8.2f} {r[11]:6.1f}")
    tot[0] += r[3]; tot[1] += r[5]; tot[2] += r[6]; tot[3] += r[8]; tot[4] += max(r[5], r[6], r[8])
print(f"# sum: measured {tot[0]:.0f} us, tensor-pipe bound {tot[1]:.0f} us, shared-memory bound {tot[2]:.0f} us, HBM bound {tot[3]:.0f} us, "
      f"per-layer max bound {tot[4]:.0f} us -> {tot[4] / tot[0]:.2f} of the measured time")
