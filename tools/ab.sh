# A/B at the shape of BASELINE config 3 with a fixed amount of work (HMC, L=10, 64 chains x 16 warps)
for i in 1 2; do
SCCODA_B200_LIB=$PWD/sccoda_b200/libA.so python tools/prof_nuts.py --method 0 --results 300 --burnin 100 | sed 's/^/A /'
python tools/prof_nuts.py --method 0 --results 300 --burnin 100 | sed 's/^/B /'
done
