#!/usr/bin/env python3
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`, no GPU needed) into markdown + a traffic JSON.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx
"""
import csv, io, json, subprocess, sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX throughput %"),
    ("lts__t_sectors_srcunit_tex_op_red.sum", "L2 RED sectors"), ("lts__t_sectors_srcunit_tex_op_atom.sum", "L2 ATOM sectors"),
    ("lts__d_atomic_input_cycles_active.max.pct_of_peak_sustained_elapsed", "L2 atomic unit busy % (max slice)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall samples: long scoreboard"),
    ("smsp__pcsamp_warps_issue_stalled_short_scoreboard", "stall samples: short scoreboard"),
    ("smsp__pcsamp_warps_issue_stalled_wait", "stall samples: wait"),
    ("smsp__pcsamp_warps_issue_stalled_lg_throttle", "stall samples: lg throttle"),
    ("smsp__pcsamp_warps_issue_stalled_mio_throttle", "stall samples: mio throttle"),
    ("smsp__pcsamp_warps_issue_stalled_barrier", "stall samples: barrier"),
    ("smsp__pcsamp_warps_issue_stalled_selected", "stall samples: selected (issuing)"),
    ("smsp__pcsamp_warps_issue_stalled_not_selected", "stall samples: not selected"),
    ("smsp__pcsamp_warps_issue_stalled_branch_resolving", "stall samples: branch resolving"),
    ("smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "stall samples: math pipe"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    md = [f"# ncu summary of `{rep.split('/')[-1]}` (ncu --set full --clock-control none)\n"]
    traffic = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0]
        md.append(f"\n## {name}\n\n| metric | value | unit |\n|---|---|---|")
        for key, label in WANT:
            if key in ix:
                md.append(f"| {label} (`{key}`) | {r[ix[key]]} | {units[ix[key]]} |")
        try:
            def to_bytes(k):
                v, u = float(r[ix[k]]), units[ix[k]].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            traffic.setdefault(name, []).append(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
        except Exception:
            pass
    open(out + ".md", "w").write("\n".join(md) + "\n")
    json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(out + "_traffic.json", "w"), indent=1)
    print("\n".join(md))


if __name__ == "__main__":
    main()
