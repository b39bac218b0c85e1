"""Top SASS instructions by warp-stall samples for one kernel of an ncu report.

    python scripts/ncu_hot.py gpurun_out/x.ncu-rep k_pinfo [top_n]
"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "sass"],
                              text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(io.StringIO(raw)))
# several launches are concatenated; take the first block
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
print(f"{kern}: {len(body)} instructions, {tot} samples")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in body) for s in stalls}
print("stall mix:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = body[i]
    st = {s[6:]: int(r[ix[s]] or 0) for s in stalls if int(r[ix[s]] or 0) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{i:5d} {int(r[ix['# Samples']]):6d}  {r[ix['Source']][:70]:70s} {st}")
