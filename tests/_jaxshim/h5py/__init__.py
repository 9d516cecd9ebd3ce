"""TEST INFRASTRUCTURE ONLY - empty stand-in so that the reference's dataloader modules import in an image without h5py; the tests
hand the loaders in-memory arrays and never open a file."""


class Dataset:
    pass


def File(*a, **k):
    raise RuntimeError('h5py stand-in of tests/_jaxshim: no HDF5 support in this image')
