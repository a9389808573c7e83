"""Turn an `ncu --set full` capture of ONE launch of the dominant kernel into the small JSON `bench.py` reads for
`roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum of that launch), keyed by kernel name, launch shape and git
revision so that a stale capture is never reported next to new timings.

usage: python tools/ncu_traffic.py REPORT.ncu-rep OUT.json KERNEL_LABEL [batch_per_gpu] [workload] [algorithmic_bytes]
  KERNEL_LABEL is what sqpqp_last_solve_kernel reports for the launch (e.g. "k_solve_cta<384,2,1>" or "k_solve_grid").
Also prints the headline counters (for profiles/*.md summaries)."""
import csv
import io
import json
import subprocess
import sys


def unit_scale(u):
    u = u.strip().lower()
    return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1.0)


def main():
    rep, out, label = sys.argv[1:4]
    batch = int(sys.argv[4]) if len(sys.argv) > 4 else None
    workload = sys.argv[5] if len(sys.argv) > 5 else None
    alg = float(sys.argv[6]) if len(sys.argv) > 6 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    best = None
    for r in rows[2:]:  # the longest launch of the report
        d = dict(zip(hdr, r))
        t = float(d.get("gpu__time_duration.sum", "0").replace(",", "") or 0)
        if best is None or t > best[0]:
            best = (t, r)
    d = dict(zip(hdr, best[1]))
    u = dict(zip(hdr, units))
    val = lambda k: float(d[k].replace(",", "")) * unit_scale(u[k]) if k in d and d[k] else None
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    rec = {"kernel": label, "ncu_kernel_name": d.get("Kernel Name"), "batch_per_gpu": batch, "workload": workload,
           "dram_bytes": (rd or 0) + (wr or 0), "dram_read": rd, "dram_write": wr,
           "duration_under_ncu": d.get("gpu__time_duration.sum") + " " + u.get("gpu__time_duration.sum", ""),
           "profile": rep.split("/")[-1], "launch": "longest launch of the capture", "algorithmic_bytes_same_launch": alg, "git": git}
    json.dump(rec, open(out, "w"), indent=1)
    print(json.dumps(rec, indent=1))
    keys = ["launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__warps_active.avg.per_cycle_active",
            "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
            "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]
    for k in keys:
        if k in d:
            print(f"{k:70s} {d[k]} {u.get(k, '')}")
    for k in hdr:
        if "issue_stalled" in k and "per_issue_active" in k:
            print(f"stall/issue {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):24s} {d[k]}")


if __name__ == "__main__":
    main()
