#!/bin/bash
# GPU box: the per-kernel ncu evidence of round 2 (profiles/kernels_*_r2.md): one `--set full` capture of every kernel of
# config 2 (all launches of one call), one per kernel class of config 3 at 1/10 scale, and the launch list of config 2.
set -u
R=${1:-r2}
mkdir -p gpurun_out
bash tools/gpu_check.sh $R launches prof_c2
bash tools/prof_each.sh $R c3s 'sort_pass_kernel' 'rank_flags_kernel<\(bool\)1>' 'rank_apply_kernel<\(bool\)1>' 'part_pass_kernel<unsigned int' 'scatter_pairs_kernel' \
    'po_group_table_kernel' 'po_scan_apply_kernel' 'po_move_kernel' 'po_round_kernel' 'po_apply_kernel' 'plcp_level_kernel' 'plcp_chunk_kernel' 'lcp_kernel' \
    'unbwt_walk' 'bwt_finish_kernel' 'pack_bytes_kernel' 'byte_hist_kernel' 'hist16_kernel' 'bucket_sort_kernel' 'part_pipe_kernel'
