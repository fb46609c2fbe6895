"""h5py stand-in: the golden generator does no file IO.  Test infrastructure only."""


class File:
    def __init__(self, *a, **k):
        raise RuntimeError("h5py is not available; the golden generator does no IO")
