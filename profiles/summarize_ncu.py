"""Compact text summary of an `ncu --set full` report (run here, no GPU needed):

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<round>_<what>.txt

One block per profiled launch with the metrics DESIGN.md / bench.py's `roofline.traffic` quote."""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2->L1 read sectors (32 B)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 LSU wavefronts %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__maximum_warps_per_active_cycle_pct", "theoretical occupancy %"),
    ("launch__occupancy_limit_registers", "occupancy limit: registers (CTAs)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit: shared memory (CTAs)"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
    ("sm__cycles_active.avg", "SM cycles active (avg)"),
]


def traffic(rep):
    """{kernel base name: DRAM bytes (read + write) per launch, averaged over the profiled launches}."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        base = "ppo_loss_kernel" if "ppo_loss_kernel" in name else "gae_scan_kernel" if "gae_scan" in name else name.split("(")[0]
        tot = sum(float(r[idx[k]]) * scale[units[idx[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        acc.setdefault(base, []).append(tot)
    return {k: sum(v) / len(v) for k, v in acc.items()}


def main():
    if sys.argv[1] == "--traffic":  # python profiles/summarize_ncu.py --traffic cfg=report.ncu-rep ... > profiles/traffic.json
        import json
        out = {}
        for arg in sys.argv[2:]:
            cfg, rep = arg.split("=", 1)
            out[cfg] = dict(traffic(rep), source=rep.split("/")[-1])
        print(json.dumps(out, indent=1))
        return
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: {len(rows) - 2} profiled launches (ncu --set full --clock-control none; cold caches, serialised)")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        print(f"\n== {name}")
        vals = {}
        for key, label in WANT:
            if key in idx:
                vals[key] = r[idx[key]]
                print(f"  {label:42s} {r[idx[key]]} {units[idx[key]]}")
        try:
            rd, wr = float(vals["dram__bytes_read.sum"]), float(vals["dram__bytes_write.sum"])
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = rd * scale[units[idx["dram__bytes_read.sum"]]] + wr * scale[units[idx["dram__bytes_write.sum"]]]
            dur = float(vals["gpu__time_duration.sum"]) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3}[units[idx["gpu__time_duration.sum"]]]
            print(f"  {'DRAM traffic (read + write)':42s} {tot / 1e6:.3f} MB -> {tot / dur / 1e9:.1f} GB/s under ncu")
        except Exception:
            pass


if __name__ == "__main__":
    main()
