"""Dense LU used by host-side chemistry setup (basis swap, logK fit).

Pure-Python restatement (IEEE doubles, same operation order) of
reference src/pflotran/utility.F90:393-476 (ludcmp) and :480-523 (lubksb):
Crout factorisation with implicit row scaling, partial pivoting where a tie
takes the LAST candidate (`>=`), and `tiny = 1e-20` on a zero pivot.
Matrices are lists of row lists, 0-based.
"""

TINY = 1.0e-20


class SingularMatrix(RuntimeError):
    pass


def ludcmp(a, n):
    indx = [0] * n
    vv = [0.0] * n
    for i in range(n):
        aamax = 0.0
        for j in range(n):
            if abs(a[i][j]) > aamax:
                aamax = abs(a[i][j])
        if aamax <= 0.0:
            raise SingularMatrix('Singular value encountered in ludcmp() row %d' % (i + 1))
        vv[i] = 1.0 / aamax
    for j in range(n):
        for i in range(j):
            s = a[i][j]
            for k in range(i):
                s = s - a[i][k] * a[k][j]
            a[i][j] = s
        aamax = 0.0
        imax = j
        for i in range(j, n):
            s = a[i][j]
            for k in range(j):
                s = s - a[i][k] * a[k][j]
            a[i][j] = s
            dum = vv[i] * abs(s)
            if dum >= aamax:
                imax = i
                aamax = dum
        if j != imax:
            a[imax], a[j] = a[j], a[imax]
            vv[imax] = vv[j]
        indx[j] = imax
        if a[j][j] == 0.0:
            a[j][j] = TINY
        if j != n - 1:
            dum = 1.0 / a[j][j]
            for i in range(j + 1, n):
                a[i][j] = a[i][j] * dum
    return indx


def lubksb(a, n, indx, b):
    ii = -1
    for i in range(n):
        ll = indx[i]
        s = b[ll]
        b[ll] = b[i]
        if ii != -1:
            for j in range(ii, i):
                s = s - a[i][j] * b[j]
        elif s != 0.0:
            ii = i
        b[i] = s
    for i in range(n - 1, -1, -1):
        s = b[i]
        for j in range(i + 1, n):
            s = s - a[i][j] * b[j]
        b[i] = s / a[i][i]
    return b
