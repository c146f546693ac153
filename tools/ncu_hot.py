"""hot instructions of one launch: by executed count and by excessive shared wavefronts
   python tools/ncu_hot.py rep launch_idx [col] [N]"""
import csv, io, subprocess, sys
path, skip = sys.argv[1], sys.argv[2]
col = sys.argv[3] if len(sys.argv) > 3 else "Instructions Executed"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and (r[idx["# Samples"]] or "0").isdigit()]
body = body[: len(body) // 2] if len(body) % 2 == 0 else body
def val(r, c):
    try: return float(r[idx[c]] or 0)
    except ValueError: return 0.0
tot = sum(val(r, col) for r in body)
print(rows[0][1][:100]); print("total", col, tot, "instructions:", len(body))
for r in sorted(body, key=lambda r: -val(r, col))[:top]:
    print(f"{val(r, col):12.0f} {100 * val(r, col) / max(tot, 1):5.1f}%  inst={val(r, 'Instructions Executed'):9.0f} shw={val(r, 'L1 Wavefronts Shared'):9.0f} ideal={val(r, 'L1 Wavefronts Shared Ideal'):9.0f}  {r[idx['Source']][:80]}")
