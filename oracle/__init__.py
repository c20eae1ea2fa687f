"""CPU oracle for the NNConv hot path -- test infrastructure only (see nnconv_oracle.py)."""
