"""Digest of the per-CTA time stamps written by scripts/halo_trace.py:   python scripts/halo_trace_report.py <dir> [pattern]"""
import glob, os, sys
import numpy as np

d = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else "trace_*.npy"
for f in sorted(glob.glob(os.path.join(d, pat))):
    a = np.load(f).astype(np.int64)            # [2][slots][ctas][8][2]
    g = a[..., 0].astype(np.float64)          # globaltimer ns
    print("==", os.path.basename(f))
    for k, name in ((0, "edge"), (1, "cell")):
        t = g[k]                               # [slots][ctas][8]
        used = t[:, :, 0] > 0
        ctas = int(used[0].sum())
        if ctas == 0:
            continue
        rows = []
        for sl in range(t.shape[0]):
            x = t[sl, :ctas]
            t0 = x[:, 0].min()
            ev = lambda e: np.where(x[:, e] > 0, x[:, e] - t0, np.nan) / 1e3       # us since the first CTA's entry
            rows.append([np.nanmax(ev(0)), np.nanmedian(ev(1)), np.nanmedian(ev(2)), np.nanmax(ev(2)), np.nanmedian(ev(3) - ev(2)), np.nanmax(ev(3) - ev(2)),
                         np.nanmedian(ev(6)), np.nanmax(ev(6)), np.nanmax(np.fmax(ev(6), ev(7)))] if k == 0 else
                        [np.nanmax(ev(0)), np.nanmedian(ev(1)), np.nanmedian(ev(3) - ev(2)) if np.isfinite(ev(2)).any() else 0.0,
                         np.nanmax(ev(3) - ev(2)) if np.isfinite(ev(2)).any() else 0.0, np.nanmin(ev(2)) if np.isfinite(ev(2)).any() else 0.0,
                         np.nanmedian(np.fmax(ev(6), ev(7))), np.nanmax(np.fmax(ev(6), ev(7))), int(np.isfinite(ev(2)).sum() + np.isfinite(ev(4)).sum()), 0.0])
        r = np.median(np.array(rows, dtype=np.float64), axis=0)
        if k == 0:
            print(f"  edge ({ctas} CTAs): last entry +{r[0]:.1f} | rows in (median) {r[1]:.1f} | first tile stored median {r[2]:.1f} max {r[3]:.1f} | "
                  f"done-count (fence) median {r[4]:.2f} max {r[5]:.2f} | CTA end median {r[6]:.1f} max {r[7]:.1f} us")
        else:
            print(f"  cell ({ctas} CTAs): last entry +{r[0]:.1f} | rows in (median) {r[1]:.1f} | flag wait median {r[2]:.2f} max {r[3]:.2f} (first waiter at {r[4]:.1f}, "
                  f"{int(r[7])} waiting groups) | groups done median {r[5]:.1f} max {r[6]:.1f} us")
    # spacing of consecutive launches (edge entry to next edge entry) = step time seen by the device
    e0 = np.sort(g[0][:, :, 0].max(axis=1))
    e0 = e0[e0 > 0]
    if e0.size > 2:
        print(f"  step spacing (edge entry to edge entry): median {np.median(np.diff(e0)) / 1e3:.2f} us")
    c0 = g[1][:, :, 0]
    # gap between the edge kernel's last CTA end and the cell kernel's first entry, and back
    ends_e = np.sort(np.nanmax(np.where(g[0][:, :, 6] > 0, g[0][:, :, 6], np.nan), axis=1))
    starts_c = np.sort(np.nanmin(np.where(c0 > 0, c0, np.nan), axis=1))
    if ends_e.size and starts_c.size:
        gaps = []
        for e in ends_e:
            nxt = starts_c[starts_c > e]
            if nxt.size:
                gaps.append(nxt[0] - e)
        if gaps:
            print(f"  edge end -> cell entry gap: median {np.median(gaps) / 1e3:.2f} us")
