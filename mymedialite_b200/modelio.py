"""MyMediaLite's text model format (IO/Model.cs:85-114, IO/MatrixExtensions.cs:31-89, IO/VectorExtensions.cs:40-60): the
layout after the two header lines is the stock classes' own, so `LoadModel` on an existing instance of the corresponding
class reads a file either side wrote (Model.cs:100-114 only warns about the type name). The header names the Cuda* class,
so the by-name route `Model.Load(filename)` (IO/Model.cs:67-83) resolves these files only where the Cuda* classes are
compiled into MyMediaLite.dll."""
import math

import numpy as np

VERSION = "2.99"


def fmt(x):
    """float.ToString(CultureInfo.InvariantCulture) on .NET Framework / Mono: 7 significant digits ("G7")."""
    x = float(np.float32(x))
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "0"
    s = "%.7G" % x
    if "E" in s:
        mant, exp = s.split("E")
        if "." in mant:
            mant = mant.rstrip("0").rstrip(".")
        sign = exp[0] if exp[0] in "+-" else "+"
        digits = exp.lstrip("+-").lstrip("0").rjust(2, "0")
        return "%sE%s%s" % (mant, sign, digits)
    return s


def write_header(w, type_name):
    w.write("%s\n%s\n" % (type_name, VERSION))


def read_header(r, expected_type):
    type_name = r.readline().rstrip("\n")
    if type_name == "":
        raise IOError("Unexpected end of file")
    r.readline()   # version line, ignored by the reference too
    return type_name


def write_vector(w, v):
    w.write("%d\n" % len(v))
    w.write("".join(fmt(x) + "\n" for x in v))


def read_vector(r):
    n = int(r.readline())
    return np.array([float(r.readline()) for _ in range(n)], np.float32)


def write_matrix(w, m):
    rows, cols = m.shape
    w.write("%d %d\n" % (rows, cols))
    for i in range(rows):
        w.write("".join("%d %d %s\n" % (i, j, fmt(m[i, j])) for j in range(cols)))
    w.write("\n")


def read_matrix(r):
    dim1, dim2 = (int(x) for x in r.readline().split(" "))
    m = np.zeros((dim1, dim2), np.float32)
    while True:
        parts = r.readline().rstrip("\n").split(" ")
        if len(parts) != 3:
            break
        i, j = int(parts[0]), int(parts[1])
        if i >= dim1:
            raise IOError("i = %d >= %d" % (i, dim1))
        if j >= dim2:
            raise IOError("j = %d >= %d" % (j, dim2))
        m[i, j] = np.float32(float(parts[2]))
    return m
