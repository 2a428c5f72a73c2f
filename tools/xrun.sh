export FPC_LIB_PATH=tools/libfpc_x.so
P="python tools/overlap_probe.py 400 600"
$P
for kb in 22 30 41 60 100; do
FPC_X_RSMEM=$kb $P
done
FPC_X_SKIPRULES=1 $P
$P
