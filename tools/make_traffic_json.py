"""Dev tool: profiles/traffic.json from the tracked ncu summaries of one measurement suite.

    python tools/make_traffic_json.py r02g          # reads profiles/r02g_ncu_full_{C4x8,C2}.csv

bench.py reads traffic.json for `roofline.traffic` (DRAM bytes per launch of the dominant kernel) and for the
warp-instruction count behind `roofline.issue_bound`, and names the csv in `roofline.traffic_source`.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = {"tri_render_fwd_kernel": "tri_render_forward", "tri_render_bwd_kernel": "tri_render_backward",
         "tri_grad_finish_kernel": "tri_grad_finish"}


def rows(path):
    with open(path) as f:
        return list(csv.DictReader(f))


def main():
    tag = sys.argv[1]
    out = {"_note": "DRAM read+write bytes and executed warp instructions PER LAUNCH of the render kernels, from "
                    "`ncu --set full` captures of tools/run_once.py (one launch each); bench.py scales them to the views "
                    "of its own launches and names the source file in roofline.traffic_source"}
    for cfg, fname, views in (("C4", "%s_ncu_full_C4x8.csv" % tag, 8), ("C2", "%s_ncu_full_C2.csv" % tag, 1)):
        rel = os.path.join("profiles", fname)
        path = os.path.join(ROOT, rel)
        if not os.path.exists(path):
            continue
        ent = {}
        for r in rows(path):
            name = r["kernel"].split("<")[0].replace("void ", "").strip()
            if name not in STAGE or STAGE[name] in ent:
                continue
            ent[STAGE[name]] = {
                "dram_bytes": int(round((float(r["dram_read"]) + float(r["dram_write"])) * 1e6)),
                "warp_instructions": int(float(r["warp_instructions"])),
                "kernel_us_under_ncu": float(r["duration_us"]),
                "views": views,
                "source": rel,
                "issue_active_pct": round(float(r["issue_active_pct"]), 1),
                "l1_data_pipe_lsu_pct": round(float(r["l1_data_pipe_lsu_pct"]), 1),
            }
        out[cfg] = ent
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
