for mode in "" "--bucketed"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 -m dml_b200.train_test --task survival --bags 48 --patches 4096 16384 --epochs 4 $mode > gpurun_out/r2_trainer_surv_varlen_2gpu$mode.log 2>&1; echo "trainer$mode exit $?"; tail -4 gpurun_out/r2_trainer_surv_varlen_2gpu$mode.log
done
