"""Turn `ncu --set full` captures (.ncu-rep, read here: no GPU needed) into the numbers the
rooflines use, so that no measurement tool carries a hand-typed constant.

    python tools/kernel_profiles.py NAME=ROLE@REP:UNITS:RAYS ... [--out profiles/r02_kernels.json]

NAME   the key of the entry: configuration + kernel ("c2_trace_stack", "c4_walk_lla_proj", ...)
ROLE   the role name of turtle_b200_kernel_info ("trace_stack", "walk_lla_proj", ...)
REP    the capture of ONE launch of that kernel
UNITS  geometry samples (trace / walk kernels) or points (query kernels) of that launch,
       as counted by the kernel itself (plan counters) and printed by the captured run
RAYS   rays / particles / points of the captured launch

Per kernel, written to the JSON (merged with what is there) and to profiles/r02_<role>.md:
  fp64_warp_inst        executed warp instructions on the FP64 pipe: the SASS source page of
                        the capture, summed over the D* opcodes (DFMA DMUL DADD DSETP DMNMX)
  fp64_lane_slots_per_sample   fp64_warp_inst x 32 / UNITS -- a warp instruction occupies 32
                        lane slots of the pipe whatever its active mask: the roofline numerator
  issue_ceiling         2 x fp64_warp_inst / all warp instructions: an FP64 warp instruction
                        holds the pipe for two issue cycles, so this is the FP64-pipe
                        utilisation at which the kernel would saturate its issue slots
  dram_bytes            dram__bytes_read.sum + dram__bytes_write.sum of the launch
  registers, duration_ms, active threads per instruction, stall breakdown ...
bench.py / tools/bench_configs.py read the JSON and refuse an entry whose `registers` differ
from the kernel they are running (turtle_b200_kernel_info).
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402

FP64_OPS = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def ncu_csv(rep, page):
    raw = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], stdout=subprocess.PIPE,
                         text=True).stdout
    return list(csv.reader([l for l in raw.splitlines() if not l.startswith("==")]))


def profile(rep, units, rays):
    rows = ncu_csv(rep, "raw")
    hdr, vals = rows[0], rows[2]
    d = dict(zip(hdr, vals))
    unit_of = dict(zip(hdr, rows[1]))

    def num(k, scale=None):
        v = float(d[k].replace(",", ""))
        u = unit_of.get(k, "")
        if scale == "bytes":
            v *= {"byte": 1., "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
        if scale == "ms":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1., "s": 1e3, "second": 1e3}[u]
        return v
    src = ncu_csv(rep, "source")
    shdr = src[1]
    i_src, i_exec = shdr.index("Source"), shdr.index("Instructions Executed")
    i_thr = shdr.index("Thread Instructions Executed")
    fp64 = total = threads = 0
    mix = {}
    for r in src[2:]:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[i_src])
        op = m.group(2) if m else "?"
        e = int(r[i_exec])
        total += e
        threads += int(r[i_thr])
        mix[op] = mix.get(op, 0) + e
        if op in FP64_OPS:
            fp64 += e
    top = sorted(mix.items(), key=lambda kv: -kv[1])[:16]
    out = {
        "kernel": d["Kernel Name"], "grid": d["Grid Size"], "block": d["Block Size"],
        "registers": int(num("launch__registers_per_thread")),
        "duration_ms": num("gpu__time_duration.sum", "ms"),
        "units": units, "rays": rays,
        "fp64_warp_inst": fp64, "warp_inst": total,
        "fp64_lane_slots_per_sample": 32. * fp64 / units,
        "warp_inst_per_sample": total / units,
        "issue_ceiling": min(1., 2. * fp64 / total),
        "active_threads_per_inst": threads / max(total, 1),
        "fp64_pipe_pct": num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "dram_bytes": num("dram__bytes_read.sum", "bytes") + num("dram__bytes_write.sum", "bytes"),
        "dram_bytes_per_sample": (num("dram__bytes_read.sum", "bytes") +
                                  num("dram__bytes_write.sum", "bytes")) / units,
        "l1_hit_pct": num("l1tex__t_sector_hit_rate.pct"),
        "l2_hit_pct": num("lts__t_sector_hit_rate.pct"),
        "local_ld_st": num("smsp__sass_inst_executed_op_local_ld.sum") +
        num("smsp__sass_inst_executed_op_local_st.sum"),
        "stall_per_issue": {k.split("stalled_")[1].split("_per_issue")[0]: num(k)
                            for k in d if k.startswith("smsp__average_warps_issue_stalled_") and
                            k.endswith("_per_issue_active.ratio") and num(k) > 0.2},
        "top_opcodes_per_sample": {k: round(v / units, 2) for k, v in top},
    }
    return out


def main():
    out_path = os.path.join(ROOT, "profiles", "r02_kernels.json")
    args = sys.argv[1:]
    if "--out" in args:
        i = args.index("--out")
        out_path = args[i + 1]
        del args[i:i + 2]
    table = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for a in args:
        name, rest = a.split("=", 1)
        role, rest = rest.split("@", 1)
        parts = rest.split(":")
        rep, units, rays = parts[0], int(float(parts[1])), int(float(parts[2]))
        entry = profile(rep, units, rays)
        entry["role"] = role
        entry["source"] = "r02_%s.md" % name
        role = name
        entry["capture"] = os.path.basename(rep)
        table[role] = entry
        md = os.path.join(ROOT, "profiles", entry["source"])
        with open(md, "w") as f:
            sys.stdout = f
            try:
                sys.argv = ["ncu_summary.py", rep]
                ncu_summary.main()
                print("## derived (tools/kernel_profiles.py)\n")
                print("| quantity | value |\n|---|---|")
                for k in ("units", "rays", "fp64_warp_inst", "warp_inst", "fp64_lane_slots_per_sample",
                          "warp_inst_per_sample", "issue_ceiling", "active_threads_per_inst",
                          "dram_bytes", "dram_bytes_per_sample"):
                    print("| %s | %s |" % (k, entry[k]))
                print("\nwarp instructions per sample, top opcodes: %s" % entry["top_opcodes_per_sample"])
            finally:
                sys.stdout = sys.__stdout__
        print("%-20s regs %3d  %.2f ms  fp64/sample %.1f  inst/sample %.1f  ceiling %.2f  threads/inst %.1f"
              % (role, entry["registers"], entry["duration_ms"], entry["fp64_lane_slots_per_sample"],
                 entry["warp_inst_per_sample"], entry["issue_ceiling"],
                 entry["active_threads_per_inst"]))
    with open(out_path, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
