"""empty stand-in: the reference imports h5py at module level but the oracle never reads .mat files"""
