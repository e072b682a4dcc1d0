#!/bin/bash
# Round 2, GPU session X: shared-memory store pattern of the K* generators (microbenchmark, with ncu wavefront counters); smoke().
mkdir -p gpurun_out
./profiles/microbench/sts_pattern > gpurun_out/sts_pattern.txt 2>&1; cat gpurun_out/sts_pattern.txt
ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,smsp__inst_executed_op_shared_st.sum --clock-control none --csv --log-file gpurun_out/sts_pattern_ncu.csv ./profiles/microbench/sts_pattern > /dev/null 2>&1
grep -v "^==" gpurun_out/sts_pattern_ncu.csv | cut -d, -f5,13,15 | head -40
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
