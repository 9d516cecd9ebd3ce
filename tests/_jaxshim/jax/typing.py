from ._array import Array as ArrayLike  # noqa: F401
