for th in 128 256; do for tile in 0 1024 1536 2048 3072 4096; do python tools/quickbench.py --configs c2_db4,c2_haar --threads $th --tile $tile --reps 20; done; done
for th in 128 256; do for fuse in 2 3 4; do python tools/quickbench.py --configs c3_sym8 --threads $th --fuse $fuse --reps 5; done; done
