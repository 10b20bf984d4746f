"""profiles/r2_traffic.json from an ncu_summarize.sh capture: DRAM bytes of the two hot kernels, tied to the
fingerprint of the kernel sources they were built from (bench.py reports roofline.traffic only when the
fingerprint matches the build it is timing).   python profiles/tools/make_traffic.py <tag> <workload> <mib>"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
tag, workload, mib = sys.argv[1], sys.argv[2], int(sys.argv[3])
m = json.load(open(os.path.join(ROOT, "gpurun_out", tag + "_metrics.json")))
units = m["units"]
def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
out = {"capture": "ncu --set full --clock-control none (profiles/tools/ncu_summarize.sh %s %s %d)" % (tag, workload, mib),
       "csrc": bench.csrc_fingerprint(), "workload": workload, "mib": mib, "kernels": {}}
for k in m["kernels"]:
    name = "aec_encode_kernel" if "aec_encode_kernel" in k["Kernel Name"] else "aec_decode_warp_kernel"
    out["kernels"][name] = {
        "kernel": k["Kernel Name"],
        "dram_bytes_read": to_bytes(k["dram__bytes_read.sum"], units.get("dram__bytes_read.sum", "byte")),
        "dram_bytes_write": to_bytes(k["dram__bytes_write.sum"], units.get("dram__bytes_write.sum", "byte")),
        "time_us": float(k["gpu__time_duration.sum"].replace(",", "")) / (1e3 if units.get("gpu__time_duration.sum") in ("ns", "nsecond") else 1),
        "alu_pipe_pct": float(k.get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "nan")),
        "fma_pipe_pct": float(k.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "nan")),
        "lsu_pct": float(k.get("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "nan")),
        "warp_instructions": float(k["smsp__inst_executed.sum"].replace(",", "")),
        "warps_active_pct": float(k.get("sm__warps_active.avg.pct_of_peak_sustained_active", "nan")),
        "registers": int(float(k.get("launch__registers_per_thread", "0"))),
        "grid": k.get("launch__grid_size"), "block": k.get("launch__block_size")}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
