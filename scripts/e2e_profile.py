"""cProfile of the end-to-end SQP solve at n=1e6 on the GPU box (host overhead hunt)."""
import cProfile, pstats, sys
sys.path[:0] = ['.', 'ip-nonlinear-solver_b200']
import bench
bench.end_to_end_solve(200000)
pr = cProfile.Profile(); pr.enable()
r = bench.end_to_end_solve(int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000)
pr.disable()
print(r)
pstats.Stats(pr).sort_stats('cumulative').print_stats(38)
