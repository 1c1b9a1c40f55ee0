# full GPU suite + bench of the current tree (one gpurun call)
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -8 gpurun_out/r2_pytest.log
timeout 1500 python -m pytest tests -m gpu -q -s -k "probe or real_dimension or config2 or vfm_head_module or postprocess" 2>&1 | grep -h "RAW label\|resize_argmax\|within-tol\|within band\|raw agreement" | cut -c1-260
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
