for d in 0 1 2 4 8 16 3 7 15 31; do echo "== dbg $d"; LIPREAD_LSTM_DBG=$d python tools/microbench.py lstm 3 2>&1 | grep "fwd_tc"; done
