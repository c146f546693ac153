"""DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels in full-set ncu captures ->
profiles/ncu_traffic.json, keyed by the C-ABI call bench.py reports.   python tools/ncu_traffic.py rep1 [rep2 ...]"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CALL_OF = {"decoder_fused_kernel": ["decoder_fused_range"], "gru_seq_tc_kernel": ["gru_seq_flagged_tc", "gru_seq_fwd_tc", "gru_seq_chunk_tc"], "gemm_tn_tc_kernel": ["gemm_tn_tc"],
           "gemm_astat_tc_kernel": ["gru_ih_gemm_tc"], "layernorm_fwd_kernel": ["layernorm_fwd"], "stft_fwd_kernel": ["stft_fwd_generic"],
           "stft512_fwd_kernel": ["stft_fwd"], "mask_istft_kernel": ["mask_istft_fwd_generic"], "mask_istft512_kernel": ["mask_istft_fwd"],
           "wo_male_partial_kernel": ["wo_male_masked_fwd"], "gru_bwd_tc_kernel": ["gru_seq_bwd_tc"]}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    acc = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        tot = sum(float(r[ix[m]]) * UNIT.get(units[ix[m]], 1.0) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        for key, calls in CALL_OF.items():
            if key in name:
                acc.setdefault(key, []).append(tot)
    for key, vals in acc.items():
        for c in CALL_OF[key]:
            out[c] = sum(vals) / len(vals)
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
old = json.load(open(path)) if os.path.exists(path) else {}
old.update(out)
json.dump(old, open(path, "w"), indent=1)
print(old)
