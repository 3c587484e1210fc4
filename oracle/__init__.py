"""TEST INFRASTRUCTURE ONLY.  CPU oracle (goi_oracle.c) and the reference's own CUDA core compiled
for sm_100a (oracle/_ref).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; nothing under goi-hyperplane_b200/ does."""
