"""amcpy.main -> amcpy_b200.main (reference: src/amcpy/main.py:160-175, console script `amcpy = amcpy.main:main`)."""
from amcpy_b200.main import cmd_extract, cmd_full, main  # noqa: F401

if __name__ == "__main__":
    main()
