import sys, cProfile, pstats, argparse
sys.path.insert(0, '/root/repo')
import bench, torch
args = argparse.Namespace(batch=0, rk_steps=100, depth=0, size=0, exchange='p2p')
w = bench.JCLindblad(args, 0, 1)
w.e2e_setup(); w.e2e_step(); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); w.e2e_step(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
