import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__shared_mem_per_block_dynamic', 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active']
extra = [h for h in hdr if 'tensor' in h.lower() or 'tmem' in h.lower() or 'utc' in h.lower()]
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print('---')
    for i in idx:
        v = r[i]
        if hdr[i] == 'Kernel Name': v = v[:70]
        print(f"  {hdr[i]}: {v} {units[i]}")
    if '-x' in sys.argv:
        for h in extra:
            print(f"  [x] {h}: {r[hdr.index(h)]} {units[hdr.index(h)]}")
