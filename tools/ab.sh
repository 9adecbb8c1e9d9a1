# A/B of two builds (libA.so = previous build) on multi-group sweeps
for i in 1; do
for c in "--conditions 3 --n-per-group 5" "--conditions 4 --n-per-group 5" "--conditions 4 --n-per-group 10" ; do
SCCODA_B200_LIB=$PWD/sccoda_b200/libA.so python tools/prof_hmc.py $c --results 200 --burnin 40 | sed 's/^/A /'
python tools/prof_hmc.py $c --results 200 --burnin 40 | sed 's/^/B /'
done
done
