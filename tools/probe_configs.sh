#!/bin/bash
# perf probes of the BASELINE.json configs other than the headline (development aid)
python tools/perf_probe.py --model funnel --data '{"D": 1}' --family sinh --chains 262144 --draws 20 --reps 3 --tag C3_funnel_sinh
python tools/perf_probe.py --model funnel --data '{"D": 10}' --family sinh --chains 262144 --draws 20 --reps 3 --tag C3_funnel11_sinh
python tools/perf_probe.py --model funnel --data '{"D": 1}' --family gauss --chains 262144 --draws 50 --reps 3 --tag funnel_gauss
python tools/perf_probe.py --model corr-normal --data '{"N": 256, "rho": 0.9}' --chains 16384 --draws 20 --reps 3 --tag C4_corr256
python tools/perf_probe.py --model arK --data @tools/ark10k.json --chains 131072 --draws 50 --reps 3 --tag C5_arK
python tools/perf_probe.py --model rosenbrock --data '{"D": 2}' --chains 262144 --draws 50 --reps 3 --tag rosenbrock
python tools/perf_probe.py --model ar1 --data '{"N": 100}' --chains 65536 --draws 100 --reps 3 --tag ar1
python tools/perf_probe.py --model normal --data '{"D": 2}' --chains 262144 --draws 200 --reps 3 --tag normal_d2
