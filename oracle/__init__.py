"""CPU oracle for the FD forward solve + adjoint.  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
The product package (red-diffeq_b200/) never imports this.
"""
