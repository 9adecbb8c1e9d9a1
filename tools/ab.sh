# A/B timing at the benchmark shape (run on the GPU box)
for i in 1 2; do
python tools/prof_hmc.py --results 400 --burnin 100 --grouped 1
done
python tools/prof_hmc.py --results 2000 --burnin 500 --grouped 1
