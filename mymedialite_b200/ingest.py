"""Host-side mirror of the reference's file readers on top of libmmlb200's native parallel parser
(csrc/ingest.cu; include/mmlb200.h "ingest").

Reference: IO/StaticRatingData.cs:36-117, IO/RatingData.cs:57-88, IO/ItemData.cs:36-93, Data/Mapping.cs:75-85,
Data/IdentityMapping.cs:62-67. Error behaviour: a malformed line raises FormatError (the reference's FormatException)
with the reference's message text.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import check

WITH_RATINGS, WITHOUT_RATINGS = "WITH_RATINGS", "WITHOUT_RATINGS"   # TestRatingFileFormat


class FormatError(ValueError):
    """System.FormatException of the reference's readers."""


class IdentityMapping:
    """Data/IdentityMapping.cs: internal id = int.Parse(original id)."""
    kind = _capi.MAP_IDENTITY

    def __init__(self):
        self.MaxEntityID = -1

    def ToOriginalID(self, internal_id):
        return str(int(internal_id))


class Mapping:
    """Data/Mapping.cs: internal ids in order of first appearance; the table lives in the last ingest that used it."""
    kind = _capi.MAP_FIRST_SEEN

    def __init__(self):
        self._table = []

    @property
    def Count(self):
        return len(self._table)

    @property
    def OriginalIDs(self):
        return list(self._table)

    def ToOriginalID(self, internal_id):
        if 0 <= internal_id < len(self._table):
            return self._table[internal_id]
        raise ValueError("Unknown internal ID: %d" % internal_id)                 # Mapping.cs:68


class ParsedFile:
    """A parsed file: COO triples in (pinned) host memory inside the library."""

    def __init__(self, h, lib):
        self.h, self.lib = h, lib
        n, mu, mi = C.c_int64(), C.c_int32(), C.c_int32()
        nu, ni, pinned = C.c_int32(), C.c_int32(), C.c_int32()
        check(lib.mml_ingest_info(h, C.byref(n), C.byref(mu), C.byref(mi), C.byref(nu), C.byref(ni), C.byref(pinned)))
        self.Count, self.MaxUserID, self.MaxItemID = n.value, mu.value, mi.value
        self.n_user_ids, self.n_item_ids, self.pinned = nu.value, ni.value, bool(pinned.value)

    def arrays(self, values=True):
        u = np.zeros(max(self.Count, 1), dtype=np.int32)
        i = np.zeros(max(self.Count, 1), dtype=np.int32)
        v = np.zeros(max(self.Count, 1), dtype=np.float32) if values else None
        check(self.lib.mml_ingest_copy(self.h, u, i, v))
        n = self.Count
        return (u[:n], i[:n], v[:n]) if values else (u[:n], i[:n])

    def original_ids(self, which):
        count = self.n_item_ids if which else self.n_user_ids
        if count == 0:
            return []
        need = C.c_int64()
        check(self.lib.mml_ingest_original_ids(self.h, which, 0, count, None, 0, None, C.byref(need)))
        buf = C.create_string_buffer(max(need.value, 1))
        off = np.zeros(count + 1, dtype=np.int64)
        check(self.lib.mml_ingest_original_ids(self.h, which, 0, count, buf, need.value, off, None))
        raw = buf.raw
        return [raw[off[j]:off[j + 1]].decode("utf-8", "replace") for j in range(count)]

    def to_device_ratings(self, ctx):
        """mml_ratings_create straight from the pinned arrays (no copy through Python)."""
        from . import engine
        h = C.c_void_p()
        check(self.lib.mml_ingest_to_ratings(ctx.h, self.h, C.byref(h)))
        r = engine.DeviceRatings.__new__(engine.DeviceRatings)
        r.ctx, r.lib, r.h = ctx, ctx.lib, h
        r.n, r.max_user, r.max_item = self.Count, self.MaxUserID, self.MaxItemID
        return r

    def to_device_feedback(self, ctx):
        from . import engine
        h = C.c_void_p()
        check(self.lib.mml_ingest_to_feedback(ctx.h, self.h, C.byref(h)))
        f = engine.DeviceFeedback.__new__(engine.DeviceFeedback)
        f.ctx, f.lib, f.h = ctx, ctx.lib, h
        f.max_user, f.max_item = self.MaxUserID, self.MaxItemID
        return f

    def close(self):
        if self.h:
            self.lib.mml_ingest_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _read(source, is_text, kind, user_mapping, item_mapping, ignore_first_line, n_threads):
    lib = _capi.load()
    user_mapping = user_mapping if user_mapping is not None else IdentityMapping()
    item_mapping = item_mapping if item_mapping is not None else IdentityMapping()
    prior = None
    for m in (user_mapping, item_mapping):
        p = getattr(m, "_last", None)
        if p is not None and p.h:
            if prior is not None and prior is not p:
                raise ValueError("user and item mappings were last used by different files")
            prior = p
    h = C.c_void_p()
    args = (kind, user_mapping.kind, item_mapping.kind, 1 if ignore_first_line else 0, int(n_threads),
            prior.h if prior is not None else None, C.byref(h))
    if is_text:
        data = source.encode("utf-8") if isinstance(source, str) else bytes(source)
        st = lib.mml_ingest_text(data, len(data), *args)
    else:
        st = lib.mml_ingest_file(str(source).encode("utf-8"), *args)
    if st == _capi.ERR_FORMAT:
        raise FormatError(lib.mml_last_error().decode("utf-8", "replace"))
    if st == _capi.ERR_IO:
        raise IOError(lib.mml_last_error().decode("utf-8", "replace"))
    check(st)
    parsed = ParsedFile(h, lib)
    for which, m in ((0, user_mapping), (1, item_mapping)):
        if isinstance(m, Mapping):
            m._table = parsed.original_ids(which)
            m._last = parsed
        else:
            m.MaxEntityID = max(m.MaxEntityID, parsed.MaxItemID if which else parsed.MaxUserID)
            m._last = parsed
    return parsed


class StaticRatingData:
    """IO/StaticRatingData.cs (and IO/RatingData.cs, whose line handling is the same)."""

    @staticmethod
    def Read(filename, user_mapping=None, item_mapping=None, test_rating_format=WITH_RATINGS, ignore_first_line=False,
             n_threads=0):
        kind = _capi.FILE_RATINGS if test_rating_format == WITH_RATINGS else _capi.FILE_RATINGS_NO_VALUE
        return _read(filename, False, kind, user_mapping, item_mapping, ignore_first_line, n_threads)

    @staticmethod
    def ReadText(text, user_mapping=None, item_mapping=None, test_rating_format=WITH_RATINGS, ignore_first_line=False,
                 n_threads=0):
        kind = _capi.FILE_RATINGS if test_rating_format == WITH_RATINGS else _capi.FILE_RATINGS_NO_VALUE
        return _read(text, True, kind, user_mapping, item_mapping, ignore_first_line, n_threads)


RatingData = StaticRatingData


class ItemData:
    """IO/ItemData.cs: implicit feedback files."""

    @staticmethod
    def Read(filename, user_mapping=None, item_mapping=None, ignore_first_line=False, n_threads=0):
        return _read(filename, False, _capi.FILE_FEEDBACK, user_mapping, item_mapping, ignore_first_line, n_threads)

    @staticmethod
    def ReadText(text, user_mapping=None, item_mapping=None, ignore_first_line=False, n_threads=0):
        return _read(text, True, _capi.FILE_FEEDBACK, user_mapping, item_mapping, ignore_first_line, n_threads)
