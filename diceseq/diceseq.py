"""`diceseq.diceseq:main` -- the entry point the reference's setup.py registers for the `diceseq` command."""
from diceseq_b200.cli import build_parser, main  # noqa: F401

if __name__ == "__main__":
    main()
