"""CPU oracle for the hot path -- test infrastructure only (see ref_numpy.py / ref_torch.py / oracle_ref.c).

Nothing under multimodaltopicsegmentation_b200/ imports this package.
"""
