"""Top stall sites of one launch in an .ncu-rep (read on the CPU box):
   python tools/ncu_stalls.py gpurun_out/x.ncu-rep <launch index> [N]"""
import csv, io, subprocess, sys


def main(path, skip, top=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--launch-skip", str(skip), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print(rows[0][1][:120])
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in rows[2:] if len(r) == len(hdr) and (r[idx["# Samples"]] or "0").isdigit()]
    tot = sum(int(r[idx["# Samples"]] or 0) for r in body)
    agg = {c: sum(int(r[idx[c]] or 0) for r in body) for c in stall_cols}
    print("total samples", tot)
    print("by reason:", {k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    body.sort(key=lambda r: -int(r[idx["# Samples"]] or 0))
    for r in body[:top]:
        n = int(r[idx["# Samples"]] or 0)
        why = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        print(f"{n:7d} {100.0 * n / max(tot, 1):5.1f}%  {r[idx['Source']][:90]:90s} {why}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 25)
