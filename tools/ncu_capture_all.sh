#!/bin/bash
# ncu --set full captures of the hot kernels on config-shaped inputs (one launch each, third call of its kind).
# Run on a GPU box:  bash tools/ncu_capture_all.sh r02 [name]   -> gpurun_out/<tag>_<name>.ncu-rep   (name: one capture only)
# (k_linear_tc matches the persistent k_linear_tcp the launcher picks; the tensor-core aggregation has its own capture,
#  tools/cfg4_agg.py under ncu -k regex:k_aggregate_tc)
tag=${1:-r01}
only=$2
cap() {  # name, kernel regex, target args...
  local name=$1 regex=$2; shift 2
  [ -n "$only" ] && [ "$only" != "$name" ] && return 0
  python tools/ncu_target.py "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --import-source on --clock-control none -k regex:$regex -s 2 -c 1 -f -o gpurun_out/${tag}_$name python tools/ncu_target.py "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -n 1 gpurun_out/ncu_$name.log
}
cap agg_cfg2 '^k_aggregate$' agg_cfg2_tri
cap aggs_cfg2 '^k_aggregate$' aggs_cfg2_tri
cap lin_fwd_fp32 k_linear_tc linear_train
cap dw_fp32 k_dw_tc linear_dw
cap epi_bwd_fp32 k_epilogue_bwd linear_epi
cap lin_fwd_bf16_h256 k_linear_tc linear_bf16_h256 2000000
