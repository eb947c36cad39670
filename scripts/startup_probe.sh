# CUDA start-up of a fresh process on this box, alone and with another process holding the GPU (under gpurun, ONE GPU)
mkdir -p gpurun_out
OUT=gpurun_out/cuda_startup.txt
nvidia-smi --query-gpu=persistence_mode,name --format=csv > $OUT
echo "# alone (no other CUDA process)" >> $OUT
for i in 1 2 3 4 5 6; do scripts/cuda_startup >> $OUT; done
echo "# CUDA_VISIBLE_DEVICES=0, alone" >> $OUT
for i in 1 2 3; do CUDA_VISIBLE_DEVICES=0 scripts/cuda_startup >> $OUT; done
scripts/cuda_startup 40 >> $OUT &
HOLD=$!
sleep 6
echo "# while another process holds a context on the GPU" >> $OUT
for i in 1 2 3 4 5 6; do scripts/cuda_startup >> $OUT; done
echo "# the same, eager module loading" >> $OUT
for i in 1 2 3; do CUDA_MODULE_LOADING=EAGER scripts/cuda_startup >> $OUT; done
kill $HOLD
cat $OUT
