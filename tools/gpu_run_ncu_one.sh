mkdir -p gpurun_out/r2; O=gpurun_out/r2
L=layer4.1.conv1
env LAYER=$L BATCH=256 WHICH=conv REPS=2 python tools/ncu_target.py > $O/plain_$L.log 2>&1 && \
env LAYER=$L BATCH=256 WHICH=conv REPS=2 ncu --set full --clock-control none --import-source on -k regex:conv_ws_kernel -s 1 -c 1 -f -o $O/ncu_$L python tools/ncu_target.py > $O/ncu_$L.log 2>&1
python tools/ncu_top.py $O/ncu_$L.ncu-rep --top 30 > $O/ncu_$L.txt 2>&1; rm -f $O/ncu_$L.ncu-rep; head -12 $O/ncu_$L.txt
