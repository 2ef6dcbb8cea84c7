"""Import-only stub: the array-level modules exercised by the golden generator never call xarray."""


class Dataset:
    pass


class DataArray:
    pass
