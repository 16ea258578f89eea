#!/bin/bash
# compute-sanitizer over the small-shape kernel tests.  One GPU.
set -u
out=gpurun_out/sanitizer
mkdir -p $out
SEL='tests/test_gpu_layers.py::test_layer_matches_reference_output tests/test_gpu_kernels.py::test_gpu_packer_bit_exact tests/test_gpu_kernels.py::test_gather_scale_bias tests/test_gpu_kernels.py::test_reference_layouts_decode_on_gpu tests/test_gpu_kernels.py::test_reference_quant3linear_module_runs_and_converts tests/test_gpu_quantize.py::test_ldlq_kernels_reproduce_the_reference_codes tests/test_gpu_quantize.py::test_ldlq_rg_with_greedy_passes_on_the_projected_layer'
SEL2="tests/test_gpu_layers.py::test_layer_token_counts tests/test_gpu_glue.py"
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --log-file $out/$tool.log \
      python -m pytest $SEL -m gpu -q -x -k "not 4096 and not 11008" > $out/${tool}_pytest.log 2>&1
  echo "$tool exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/$tool.log | tail -1)" | tee -a $out/summary.txt
  tail -2 $out/${tool}_pytest.log
done
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $out/memcheck2.log \
    python -m pytest $SEL2 -m gpu -q -x -k "not 2048" > $out/memcheck2_pytest.log 2>&1
echo "memcheck (token counts, glue) exit $? : $(grep -E 'ERROR SUMMARY' $out/memcheck2.log | tail -1)" | tee -a $out/summary.txt
tail -2 $out/memcheck2_pytest.log
