#!/bin/bash
# round 2, GPU call V: compute-sanitizer on the round-2 kernels (memcheck on the parity / power / synth suites, racecheck and
# initcheck on the cases that run the shared-memory kernels: column-blocked long rows, band kernels, SM-affine queues)
OUT=gpurun_out/r02v
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
( time timeout 1500 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_power.py tests/test_gpu_synth.py -m gpu -q -x --timeout 1400 -p no:cacheprovider -k "not reference_main and not c_example and not sm_affine" ) > $OUT/memcheck.log 2>&1
echo "memcheck rc=$?" >> $OUT/memcheck.log; tail -8 $OUT/memcheck.log
( time timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 800 -p no:cacheprovider -k "test_preprocessing_bit_exact_and_spmv and (skewed or powerlaw or mixed or one_row or long_pad)" ) > $OUT/racecheck.log 2>&1
echo "racecheck rc=$?" >> $OUT/racecheck.log; tail -6 $OUT/racecheck.log
( time timeout 900 $CS --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 800 -p no:cacheprovider -k "test_preprocessing_bit_exact_and_spmv and (skewed or powerlaw or mixed or wide_span)" ) > $OUT/initcheck.log 2>&1
echo "initcheck rc=$?" >> $OUT/initcheck.log; tail -6 $OUT/initcheck.log
( time DASP_SMQ=1 timeout 600 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 500 -p no:cacheprovider -k "test_preprocessing_bit_exact_and_spmv and (mixed or powerlaw or ragged)" ) > $OUT/memcheck_smq.log 2>&1
echo "memcheck smq rc=$?" >> $OUT/memcheck_smq.log; tail -5 $OUT/memcheck_smq.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 500 -p no:cacheprovider -k "sm_affine" > $OUT/pytest_smq.log 2>&1; tail -3 $OUT/pytest_smq.log
echo done
