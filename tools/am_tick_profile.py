"""cProfile of one C4-scale Alonso-Mora tick (see tools/am_tick_c4.py): where the host time goes."""
import cProfile
import io
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import am_tick_c4 as A  # noqa: E402


class Args:
    vehs, reqs = 10000, 2000


def main():
    cap = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    G, vehs, reqs, T, _ = A.build(Args)
    A.one_cap(G, vehs, reqs, T, cap, "/dev/null")
    pr = cProfile.Profile()
    pr.enable()
    A.one_cap(G, vehs, reqs, T, cap, "/dev/null")
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
    print(s.getvalue())


if __name__ == "__main__":
    main()
