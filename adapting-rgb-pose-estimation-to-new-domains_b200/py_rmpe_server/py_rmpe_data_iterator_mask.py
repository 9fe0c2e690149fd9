"""The reference keeps a second copy of the iterator in py_rmpe_server/py_rmpe_data_iterator_mask.py (same class, same
signatures); here both module paths resolve to the one implementation in py_rmpe_data_iterator.py."""
from .py_rmpe_data_iterator import RawDataIterator  # noqa: F401
