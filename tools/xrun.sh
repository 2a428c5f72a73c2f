export FPC_LIB_PATH=tools/libfpc_x.so
P="python tools/overlap_probe.py 400 600"
$P
for us in 3 6 10 15 20 30; do
FPC_X_RDELAY=$us $P
done
$P
