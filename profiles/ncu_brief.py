"""One line per kernel launch of an .ncu-rep: duration, DRAM bytes, registers, warps/SM, lanes per instruction, warp
instructions, issue-active %, shared-memory wavefronts / bank conflicts, L2 hit rate, DRAM throughput %."""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, scale=1.0):
    try:
        return float(r[col[name]].replace(",", "")) * scale
    except (KeyError, ValueError):
        return float("nan")


def to_unit(r, name, want):
    v, u = val(r, name), units[col[name]] if name in col else ""
    f = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    return v * f


print(f"{'kernel':34s} {'us':>9s} {'rdMB':>8s} {'wrMB':>8s} {'regs':>4s} {'warps%':>6s} {'thr/i':>5s} {'Minst':>8s} {'issue%':>6s} "
      f"{'smemMwf':>8s} {'conflM':>7s} {'L2hit%':>6s} {'dram%':>6s}")
for r in rows[2:]:
    print(f"{r[col['Kernel Name']].split('(')[0][-34:]:34s} {to_unit(r, 'gpu__time_duration.sum', 'us'):9.1f} "
          f"{to_unit(r, 'dram__bytes_read.sum', 'MB'):8.1f} {to_unit(r, 'dram__bytes_write.sum', 'MB'):8.1f} "
          f"{val(r, 'launch__registers_per_thread'):4.0f} {val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{val(r, 'smsp__thread_inst_executed_per_inst_executed.ratio'):5.1f} {val(r, 'smsp__inst_executed.sum', 1e-6):8.2f} "
          f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{val(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 1e-6):8.2f} "
          f"{val(r, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 1e-6):7.2f} "
          f"{val(r, 'lts__t_sector_hit_rate.pct'):6.1f} {val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f}")
