export FPC_LIB_PATH=tools/libfpc_x.so
export FPC_P_NOPROF=1
FPC_P_SYNC=1 python tools/overlap_probe.py 200 600
python tools/overlap_probe.py 20 600
python tools/overlap_probe.py 400 600
FPC_P_CPUDELAY=40 python tools/overlap_probe.py 200 600
FPC_P_CPUDELAY=70 python tools/overlap_probe.py 200 600
