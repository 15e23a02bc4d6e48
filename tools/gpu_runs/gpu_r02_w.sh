mkdir -p gpurun_out
timeout 600 python tools/ab_cls_last.py > gpurun_out/w_ab_cls.txt 2>&1
cat gpurun_out/w_ab_cls.txt
