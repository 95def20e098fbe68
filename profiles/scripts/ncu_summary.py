"""Write profiles/kernel_metrics.json from `ncu --set full` captures (.ncu-rep) of the bench-shaped step (prof_r2_checkpoint.sh, verify_and_capture.sh):
per kernel the DRAM bytes per launch and the utilisation numbers bench.py quotes in `roofline`.  Keyed by the sha1 of
orb-ygz-slam_b200/csrc (bench.csrc_sha) so a stale capture is never quoted for a newer build.
usage: python profiles/scripts/ncu_summary.py BATCH out.json rep1.ncu-rep [rep2 ...]"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench

STAGE = {"k_fast_cells_v2": "fast_cells", "k_pyr_resize_v2": "pyramid(7 launches)", "k_octree": "octree", "k_blur_levels": "blur_levels",
         "k_orient_describe": "orient_describe", "k_bow_match_batch": "bow_match", "k_vocab_transform": "vocab_transform"}
M = {"dram_bytes_read": "dram__bytes_read.sum", "dram_bytes_write": "dram__bytes_write.sum", "duration_ns": "gpu__time_duration.sum",
     "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
     "alu_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "fma_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
     "lsu_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "warp_instructions": "smsp__inst_executed.sum",
     "registers": "launch__registers_per_thread"}
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1, "usecond": 1e3, "msecond": 1e6, "nsecond": 1, "us": 1e3, "ms": 1e6, "ns": 1}

batch, out = int(sys.argv[1]), sys.argv[2]
kernels = {}
for rep in sys.argv[3:]:
    # a .ncu-rep is read through ncu; a .csv is the raw page already exported (ncu -i rep --page raw --csv)
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        name = r[kn].split("(")[0].split("::")[-1].split("<")[0]          # k_octree<4>, k_vocab_transform<10> -> base name
        st = STAGE.get(name)
        if not st:
            continue
        d = kernels.setdefault(st, {"kernel": name, "launches": 0})
        d["launches"] += 1
        for k, col in M.items():
            if col in hdr:
                i = hdr.index(col)
                v = float(r[i].replace(",", "")) * UNIT.get(units[i], 1)
                d[k] = d.get(k, 0) + v if k in ("dram_bytes_read", "dram_bytes_write", "duration_ns", "warp_instructions") else max(d.get(k, 0), v)
json.dump({"csrc_sha": bench.csrc_sha(), "batch": batch, "source": "ncu --set full --clock-control none, " + ", ".join(os.path.basename(x) for x in sys.argv[3:]),
           "note": "sums over the launches of a stage (7 pyramid levels); utilisation = max over them", "kernels": kernels}, open(out, "w"), indent=1)
print(json.dumps(kernels, indent=1))
