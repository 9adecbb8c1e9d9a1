for i in 1 2; do
SCCODA_B200_LIB=$PWD/sccoda_b200/libA.so python tools/prof_hmc.py --results 400 --burnin 100 | sed 's/^/A /'
python tools/prof_hmc.py --results 400 --burnin 100 | sed 's/^/B /'
done
