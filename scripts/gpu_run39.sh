mkdir -p gpurun_out
SSDE_LIB_SUFFIX=_sp1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "test_nllk_and_gradient_match_oracle and CTCRW and not CTCRW-5" 2>&1 | tail -3
bash scripts/gpu_tune.sh t7 rs1 sp1 rs1 sp1
