cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "search or edge or encode_matches or large_batch or trained" > gpurun_out/r2_tests2.log 2>&1; echo "tests rc=$?"
python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "interleave or tail or product or stats" > gpurun_out/r2_tests2b.log 2>&1; echo "tests2b rc=$?"
nvcc -o /tmp/san_search tools/san_search.cu -Lquantization_b200 -lmcq -Xlinker -rpath=$PWD/quantization_b200 > gpurun_out/r2_san.log 2>&1
for mode in 0 1 2; do for n in 2 4 8 16; do /tmp/san_search $n 512 $mode >> gpurun_out/r2_san.log 2>&1; echo "san n=$n mode=$mode rc=$?" >> gpurun_out/r2_san.log; done; done
for v in "" _nopair _u8 _u2 _nopair_u8; do
  echo "== variant libmcq$v" >> gpurun_out/r2_bsearch2.log
  MCQ_ONLY=v2 MCQ_LIB=$PWD/quantization_b200/libmcq$v.so python tools/bench_search.py 75776 8 512 >> gpurun_out/r2_bsearch2.log 2>&1
done
echo "== N=16 D=1024" >> gpurun_out/r2_bsearch2.log
MCQ_ONLY=v2 python tools/bench_search.py 37888 16 1024 >> gpurun_out/r2_bsearch2.log 2>&1
echo "== N=4 D=256" >> gpurun_out/r2_bsearch2.log
MCQ_ONLY=v2 python tools/bench_search.py 75776 4 256 >> gpurun_out/r2_bsearch2.log 2>&1
tail -3 gpurun_out/r2_tests2.log; tail -3 gpurun_out/r2_tests2b.log; grep -c "rc=0" gpurun_out/r2_san.log; cat gpurun_out/r2_bsearch2.log | grep -v "^frame_passes"
