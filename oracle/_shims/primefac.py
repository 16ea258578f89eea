"""Stand-in for the third-party `primefac` package (not installed, no network).

The reference uses exactly one entry point, `primefac.primefac(n)`, as an
iterator of the prime factors of n in ascending order (method.py:16-18).
Only `oracle/gen_golden.py` puts this directory on sys.path, to import the
reference in the build container.
"""


def primefac(n):
    n = int(n)
    d = 2
    while d * d <= n:
        while n % d == 0:
            yield d
            n //= d
        d += 1 if d == 2 else 2
    if n > 1:
        yield n
