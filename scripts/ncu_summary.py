"""Text summary of an ncu report: ncu -i <rep> --page raw --csv | python scripts/ncu_summary.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
cols = {"dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "dram%": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "tensor%": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "regs": "launch__registers_per_thread", "grid": "launch__grid_size", "block": "launch__block_size", "smem": "launch__shared_mem_per_block_dynamic",
        "warps%": "sm__warps_active.avg.pct_of_peak_sustained_active", "sm_clk": "sm__cycles_elapsed.avg.per_second"}
idx = {k: hdr.index(v) for k, v in cols.items() if v in hdr}
name_i = hdr.index("Kernel Name")
def scale(u):
    u = u.lower()
    return {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1, "second": 1, "nsecond": 1e-9}.get(u, 1)
print(f"{'kernel':46s} {'dur us':>9s} {'DRAM rd MB':>10s} {'wr MB':>8s} {'GB/s':>7s} {'dram%':>6s} {'tensor%':>7s} {'regs':>4s} {'grid':>6s} {'blk':>4s} {'smem KB':>7s}")
for r in rows[2:]:
    g = lambda k: float(r[idx[k]].replace(",", "")) if k in idx and r[idx[k]] not in ("", "n/a") else float("nan")
    dur = g("dur") * scale(units[idx["dur"]]); rd = g("rd") * scale(units[idx["rd"]]); wr = g("wr") * scale(units[idx["wr"]])
    name = r[name_i].replace("void ", "").replace("<unnamed>::", "")[:46]
    print(f"{name:46s} {dur * 1e6:9.1f} {rd / 1e6:10.1f} {wr / 1e6:8.1f} {(rd + wr) / dur / 1e9:7.0f} {g('dram%'):6.1f} {g('tensor%'):7.1f} {int(g('regs')):4d} {int(g('grid')):6d} {int(g('block')):4d} {g('smem') * scale(units[idx['smem']]) / 1e3 if 'smem' in idx else float('nan'):7.1f}")
