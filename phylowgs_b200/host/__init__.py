"""Python-3 host glue around the CUDA library: the reference's tssb.py / params.py / evolve.py / multievolve.py interfaces with
the likelihood-bound calls routed to libpwgs_b200.so.  No likelihood arithmetic lives here."""
