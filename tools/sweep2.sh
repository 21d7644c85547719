for fuse in 1 2 3 4; do python tools/quickbench.py --configs c4_coif5 --fuse $fuse --reps 3; done
for th in 128 256; do python tools/quickbench.py --configs c4_coif5 --threads $th --reps 3; done
python tools/quickbench.py --configs c5_db8 --mode 1 --reps 3
python tools/quickbench.py --configs c5_db8 --mode 2 --reps 3
python tools/quickbench.py --configs c2s_db4 --reps 50
