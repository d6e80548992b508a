python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for w in 1 4 8; do SC2_CODER_WARPS=$w python scripts/diag_coder.py 3 256 | tail -1; done
python scripts/diag_coder.py 3 1 | tail -1
for w in 4 8; do SC2_CODER_WARPS=$w python scripts/diag_pipeline.py 8 96 | tail -1; done
