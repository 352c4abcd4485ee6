"""python train.py --cfg cfg/p16t9c85r12.cfg --band NIR   (the reference's training entry point, on the B200 engine)."""
from probav_b200.cli import train_main

if __name__ == "__main__":
    train_main()
