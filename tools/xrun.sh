export FPC_LIB_PATH=tools/libfpc_x.so
P="python tools/overlap_probe.py 300 600"
for parts in 1 2 4 7; do
FPC_X_PARTS=$parts $P
FPC_X_PARTS=$parts FPC_X_SKIPRULES=1 $P
done
