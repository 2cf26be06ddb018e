for lib in "$@"; do export DVBS2B200_LIB=$PWD/gr-dvbs2rx_b200/$lib; echo $lib;
python tools/run_one.py fec C1_2 1 1.0 25 2664 2>&1 | tail -1;
python tools/run_one.py fec C1_2 1 2.0 25 2664 2>&1 | tail -1;
python tools/run_one.py fec C3_4 1 4.6 50 2664 2>&1 | tail -1;
python tools/run_one.py fec C3_5 1 3.5 25 2664 2>&1 | tail -1;
python tools/run_one.py fec C9_10 1 6.6 25 2664 2>&1 | tail -1;
done
