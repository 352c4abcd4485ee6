"""python test.py --cfg cfg/p16t9c85r12.cfg --band NIR --totest TEST   (the reference's predict entry point, on the B200 engine)."""
from probav_b200.cli import test_main

if __name__ == "__main__":
    test_main()
