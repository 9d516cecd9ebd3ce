"""TEST INFRASTRUCTURE ONLY - empty stand-in so that the reference's dataloader modules import in an image without imageio."""


class _FreeImage:
    @staticmethod
    def download():          # dsec_loader.py:14 calls it at import time (a network fetch in the real package)
        return None


class plugins:
    freeimage = _FreeImage
