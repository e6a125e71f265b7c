"""``python -m sisi4s_b200 in.yaml`` -- run a sisi4s-style execution plan (see plan.py)."""
import sys

from .plan import run_plan_file


def main(argv):
    if len(argv) != 2 or argv[1] in ("-h", "--help"):
        print(__doc__)
        return 2
    run_plan_file(argv[1])
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
