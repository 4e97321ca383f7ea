for nb in 128 256; do echo "== n=2000 nb=$nb"; python tools/batch_bench.py --sizes 1000,2000,3000 --nb $nb --reps 4 2>&1 | cut -c1-360; done
for nb in 128 256 384 512; do echo "== n=5000 nb=$nb"; python tools/batch_bench.py --sizes 5000 --nb $nb --reps 3 2>&1 | cut -c1-360; done
for nb in 256 512; do echo "== n=10000 nb=$nb"; python tools/batch_bench.py --sizes 10000 --nb $nb --reps 2 2>&1 | cut -c1-360; done
for r in 8 16 24 40; do echo "== n=2000 VC_RESERVE_SMS=$r"; VC_RESERVE_SMS=$r python tools/batch_bench.py --sizes 2000,5000 --reps 3 2>&1 | cut -c1-360; done
