export FPC_LIB_PATH=tools/libfpc_x.so
export FPC_P_NOPROF=1
for k in 10 20 40 100 400; do
python tools/overlap_probe.py $k 600
FPC_NO_STARTUP_STAMPS=1 python tools/overlap_probe.py $k 600
done
