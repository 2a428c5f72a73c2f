export FPC_LIB_PATH=tools/libfpc_x.so
P="python tools/overlap_probe.py 400 600"
$P
FPC_X_NOCOUNTERS=1 $P
$P
FPC_X_NOCOUNTERS=1 $P
