set -x
python tests/notes/win_ncu.py 1024 > gpurun_out/r2_win_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:conv_up_win -s 6 -c 3 -o gpurun_out/r2_prof_win -f python tests/notes/win_ncu.py 1024 > gpurun_out/r2_ncu_win.log 2>&1
tail -3 gpurun_out/r2_ncu_win.log
