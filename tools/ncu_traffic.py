#!/usr/bin/env python
"""Per-operation DRAM traffic from an `ncu --set full ... --page raw --csv` export of one training step:
python tools/ncu_traffic.py raw.csv out.json "capture description".
An operation is identified by its kernel instantiation; where two layers share one (e.g. convt5.fwd and
conv1.dgrad are both tc2_kernel<8,1,1>) the decoder's launch is the longer one (9x the images)."""
import csv, json, re, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
unit = {n: rows[1][i] for n, i in col.items()}


def val(r, name):
    v = float(r[col[name]].replace(",", ""))
    u = unit[name]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3,
                "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}.get(u, 1.0)


OPS = [  # op, kernel regex, pick ("max": longest launch, "min": shortest) — round-2 instantiations <CIN, COUT, SD, TMA>
    ("convt5.fwd", r"tc2_kernel<8, 1, 1, 1>", "max"), ("conv1.dgrad", r"tc2_kernel<8, 1, 1, 0>", "max"),
    ("convt5.dgrad", r"tc2_kernel<1, 8, 1, 0>", "max"), ("conv1.fwd", r"gather_kernel<1, 8, 4>", "max"),
    ("convt4.fwd", r"tc2_kernel<8, 8, 1, 1>", "max"), ("conv2.dgrad", r"tc2_kernel<8, 8, 1, 0>", "max"),
    ("convt4.dgrad", r"tc2_kernel<8, 8, 2, 1>", "max"), ("convt2.dgrad", r"tc2_kernel<16, 16, 2, 0>", "max"),
    ("convt3.fwd", r"tc2_kernel<16, 8, 1, 0>", "max"), ("convt3.dgrad", r"tc2_kernel<8, 16, 1, 0>", "max"),
    ("convt5.wgrad", r"wgrad_mma_kernel<1, 8, 2, 1>", "max"), ("convt4.wgrad", r"wgrad_mma_kernel<8, 8, 12, 1>", "max"),
    ("convt3.wgrad", r"wgrad_mma_kernel<8, 16, 7, 1>", "max"), ("convt2.wgrad", r"wgrad_mma_kernel<16, 16, 7, 1>", "max"),
    ("conv2.fwd", r"gather_kernel<8, 8, 4>", "max"), ("conv3.fwd", r"gather_kernel<8, 16, 4>", "max"),
    ("bnt3.bn_bwd", r"bn_bwd_apply_kernel<16>", "max"), ("convt5.box_sums", r"box_sums_kernel", "max"),
    ("recon_loss.fwd", r"recon_fwd_kernel", "max"), ("recon_loss.bwd", r"recon_bwd_kernel", "max"),
    ("fc2-fc43.fwd", r"mlp_fwd_kernel<1>", "max"), ("fc5-fc7.fwd", r"mlp_fwd_kernel<2>", "max"),
    ("fc2-fc43.bwd", r"mlp_bwd_kernel<1>", "max"), ("fc5-fc7.bwd", r"mlp_bwd_kernel<2>", "max"),
]
out = {"capture": sys.argv[3] if len(sys.argv) > 3 else "", "ops": {}}
for op, rx, pick in OPS:
    cand = [r for r in data if re.search(rx, r[col["Kernel Name"]])]
    if not cand:
        continue
    key = lambda r: val(r, "gpu__time_duration.sum")
    r = max(cand, key=key) if pick == "max" else min(cand, key=key)
    tp = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    out["ops"][op] = {"kernel": r[col["Kernel Name"]][:70], "dram_read_bytes": val(r, "dram__bytes_read.sum"),
                      "dram_write_bytes": val(r, "dram__bytes_write.sum"),
                      "ncu_duration_us": round(val(r, "gpu__time_duration.sum"), 2),
                      "tensor_pipe_pct": float(r[col[tp]]) if tp in col else None}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out["ops"], indent=1)[:1500])
