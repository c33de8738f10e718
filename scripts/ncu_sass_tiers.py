#!/usr/bin/env python
"""Dev: splits the SASS-level stall samples of the panel-GEMM function of an .ncu-rep (source page) into execution-count
tiers (k-step loop / per chunk / per round / per call) and lists the hottest instructions outside the k-step loop."""
import csv, collections, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) > 10 and r[0].startswith("0x")]
tot = sum(int(r[2]) for r in data)
ublk = [i for i, r in enumerate(data) if "UBLKCP" in r[1]]
dm = [i for i, r in enumerate(data) if "DMMA" in r[1] and i > ublk[0] - 3000]
# the function: from the BAR/first instruction after the previous RET/EXIT to the RET after the last DMMA
prev = [i for i, r in enumerate(data) if ("EXIT" in r[1] or "RET" in r[1]) and i < ublk[0]]
nxt = [i for i, r in enumerate(data) if "RET" in r[1] and i > max(dm)]
lo, hi = prev[-1] + 1, nxt[0]
fn = data[lo:hi + 1]
s = sum(int(r[2]) for r in fn)
print(f"panel function: SASS rows {lo}..{hi}, {100 * s / tot:.1f} % of all samples")
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
st = {h: sum(int(r[ix[h]] or 0) for r in fn) for h in stall}
print("  stalls:", ", ".join(f"{k[6:]} {100 * v / s:.1f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
mx = max(int(r[5]) for r in fn if "DMMA" in r[1])
tiers = collections.OrderedDict((k, [0, 0]) for k in ("kstep", "chunk", "round", "call"))
def tier(e): return "kstep" if e > 0.6 * mx else "chunk" if e > 0.12 * mx else "round" if e > 0.004 * mx else "call"
for r in fn:
    t = tiers[tier(int(r[5]))]; t[0] += int(r[2]); t[1] += 1
for k, (v, n) in tiers.items(): print(f"  tier {k:6s}: {100 * v / s:5.1f} % of the function's samples ({100 * v / tot:4.1f} % of all), {n} instructions")
dmma = sum(int(r[2]) for r in fn if "DMMA" in r[1]); nop = sum(int(r[2]) for r in fn if r[1].strip().endswith("NOP") and tier(int(r[5])) == "kstep")
print(f"  DMMA + NOP samples in the k-step loop: {100 * (dmma + nop) / s:.1f} % of the function")
for name in ("chunk", "round"):
    print(f"  hottest {name}-tier instructions:")
    for r in sorted((r for r in fn if tier(int(r[5])) == name), key=lambda r: -int(r[2]))[:12]:
        why = ", ".join(f"{h[6:]}:{r[ix[h]]}" for h in stall if r[ix[h]] not in ("0", "") and int(r[ix[h]]) > int(r[2]) * 0.3)
        print(f"     {r[1].strip()[:58]:58s} {int(r[2]):8d} x{int(r[5]):10d}  {why}")
