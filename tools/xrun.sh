export FPC_LIB_PATH=tools/libfpc_x.so
export FPC_P_NOPROF=1
export FPC_NO_STARTUP_STAMPS=1
echo "plain events, legacy stream: K=20, K=400, SYNC"
python tools/overlap_probe.py 20 600
python tools/overlap_probe.py 400 600
FPC_P_SYNC=1 python tools/overlap_probe.py 200 600
echo "plain events, own stream: K=20, K=400, SYNC"
export FPC_P_OWNSTREAM=1
python tools/overlap_probe.py 20 600
python tools/overlap_probe.py 400 600
FPC_P_SYNC=1 python tools/overlap_probe.py 200 600
