# ncu_summarize.sh <tag> <workload> <mib>: full capture of the encode and warp-decode kernels of one workload,
# reduced ON THE BOX to small text files (raw metrics per kernel, per-source-line instruction and stall
# shares); the .ncu-rep itself is deleted (gpurun_out/ carries at most 64 MiB back)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; T=$1; W=$2; M=$3
REP=/tmp/prof_$T
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"aec_(encode|decode_warp)_kernel" -s 6 -c 2 -f -o $REP python bench.py --device-only --steps 1 --warmup 3 --workload $W --mib $M > gpurun_out/${T}_ncu.log 2>&1
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/${T}_raw.csv 2>/dev/null
python - $T <<'PY'
import csv, sys, json
T = sys.argv[1]
rows = list(csv.reader(open("gpurun_out/%s_raw.csv" % T)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed.sum", "smsp__average_warp_latency_issue_stalled_barrier.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
out = []
for r in rows[2:]:
    d = {}
    for w in want:
        if w in hdr:
            d[w] = r[hdr.index(w)]
    # every stall reason column
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") is False and "warps_issue_stalled" in h and h.endswith(".pct"):
            d[h] = r[i]
    out.append(d)
json.dump({"units": dict(zip(hdr, rows[1])) if len(rows) > 1 else {}, "kernels": out}, open("gpurun_out/%s_metrics.json" % T, "w"), indent=1)
for d in out:
    print({k: d[k] for k in list(d)[:14]})
PY
mkdir -p /tmp/dis_$T && cd /tmp/dis_$T && cuobjdump -xelf all $GRAFT_REPO_ROOT/libaec_b200/lib/libaec.so.0 > /dev/null && nvdisasm -g -c aec_encode.sm_100a.cubin > enc_dis.txt 2>/dev/null; nvdisasm -g -c aec_decode.sm_100a.cubin > dec_dis.txt 2>/dev/null; cd $GRAFT_REPO_ROOT
ENCK=$(grep -o 'aec_encode_kernel<[0-9]*, [0-9]*>' gpurun_out/${T}_raw.csv | head -1); DECK=$(grep -o 'aec_decode_warp_kernel<[0-9]*, [0-9]*>' gpurun_out/${T}_raw.csv | head -1)
EJ=$(echo $ENCK | sed 's/.*<\([0-9]*\), \([0-9]*\)>/ILi\1ELi\2E/'); DJ=$(echo $DECK | sed 's/.*<\([0-9]*\), \([0-9]*\)>/ILi\1ELi\2E/')
ncu -i $REP.ncu-rep --page source --csv --kernel-name regex:aec_encode_kernel > /tmp/enc_sass_$T.csv 2>/dev/null
ncu -i $REP.ncu-rep --page source --csv --kernel-name regex:aec_decode_warp_kernel > /tmp/dec_sass_$T.csv 2>/dev/null
python profiles/tools/ncu_by_line.py /tmp/enc_sass_$T.csv /tmp/dis_$T/enc_dis.txt "aec_encode_kernel$EJ" 70 > gpurun_out/${T}_encode_by_line.txt 2>&1
python profiles/tools/ncu_by_line.py /tmp/dec_sass_$T.csv /tmp/dis_$T/dec_dis.txt "aec_decode_warp_kernel$DJ" 70 > gpurun_out/${T}_decode_by_line.txt 2>&1
head -5 gpurun_out/${T}_encode_by_line.txt; head -5 gpurun_out/${T}_decode_by_line.txt
rm -f $REP.ncu-rep gpurun_out/${T}_raw.csv.bak
