"""Condense an .ncu-rep into the few numbers DESIGN.md / bench.py quote:  python tools/ncu_summary.py rep.ncu-rep out.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("smsp__cycles_elapsed.avg.per_second", "SM clock"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/CTA"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data-pipe wavefronts % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "shared-atomic wavefronts"),
    ("smsp__inst_executed_op_shared_atom.sum", "shared-atomic warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "shared-load wavefronts"),
    ("l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed", "TEX data-pipe wavefronts % of peak"),
    ("l1tex__data_pipe_tex_wavefronts_mem_texture.sum", "texture wavefronts"),
    ("l1tex__t_sector_pipe_tex_mem_texture_hit_rate.pct", "texture L1 sector hit rate %"),
    ("smsp__inst_executed_pipe_tex.sum", "texture-pipe warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum", "shared-atomic bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# ncu summary of `%s`\n\n" % rep.split("/")[-1])
        f.write("Captured with `ncu --set full --clock-control none --import-source on` (one GPU). Per-launch values.\n\n")
        names = [r[idx["Kernel Name"]] for r in data]
        f.write("| metric | " + " | ".join("launch %d" % i for i in range(len(data))) + " | unit |\n")
        f.write("|---|" + "---|" * (len(data) + 1) + "\n")
        f.write("| kernel | " + " | ".join(n.replace("|", "/")[:60] for n in names) + " | |\n")
        for key, label in KEYS:
            if key in idx:
                f.write("| %s (`%s`) | %s | %s |\n" % (label, key, " | ".join(r[idx[key]] for r in data), units[idx[key]]))
        tr = []
        for r in data:
            try:
                a, b = float(r[idx["dram__bytes_read.sum"]]), float(r[idx["dram__bytes_write.sum"]])
                ua = units[idx["dram__bytes_read.sum"]]
                mul = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(ua, 1)
                tr.append("%.4g" % ((a + b) * mul))
            except Exception:
                tr.append("n/a")
        f.write("| **traffic = read + write** | %s | byte |\n" % " | ".join(tr))


if __name__ == "__main__":
    main()
