for w in teapot1080 sweep1080; do for n in 1 2 3 4 5 6 7 0; do echo -n "$w stop=$n "; RXC_FRONT_STOP=$n python tools/quick_bench.py $w 2>&1 | grep -v Warn | tail -1 | awk '{print $NF}'; done; done
