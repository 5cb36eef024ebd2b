for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu -k back_to_back 2>&1 | tail -1; done
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
