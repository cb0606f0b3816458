"""python -m ribodetector_b200 ... = the `ribodetector` command line."""
from .detect import main

if __name__ == "__main__":
    main()
