"""Headless no-op stand-in for ``pygame`` -- TEST INFRASTRUCTURE ONLY.

The reference's world model imports pygame at module scope
(/root/reference/cosmosim/core/universe_old.py:1, planet.py:1) and pygame is not
installed in this image.  ``oracle/gen_golden.py`` puts this directory on
``sys.path`` so that the *unmodified* reference ``Universe.simulate()`` can be
driven headlessly through its own ``run_while`` hook.  Nothing here draws or
reads input; every call the frame loop makes returns an inert object.
"""

QUIT, KEYDOWN, MOUSEBUTTONDOWN, MOUSEBUTTONUP, MOUSEMOTION = 256, 768, 1025, 1026, 1024
K_SPACE, K_w, K_a, K_s, K_d, K_KP_PLUS, K_KP_MINUS = 32, 119, 97, 115, 100, 1073741911, 1073741910


class _Inert:
    """Object whose every attribute is a callable returning another _Inert."""

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        return _Inert()

    def __iter__(self):
        return iter(())

    def __int__(self):
        return 0

    def __float__(self):
        return 0.0


class _Clock:
    def tick(self, fps=0):
        return 0

    def get_time(self):
        return 1

    def get_fps(self):
        return 0.0


class _Time:
    Clock = _Clock


class _Event:
    @staticmethod
    def get():
        return []


class _Mouse:
    @staticmethod
    def get_pos():
        return (0, 0)


time = _Time()
event = _Event()
mouse = _Mouse()
font = _Inert()
display = _Inert()
draw = _Inert()


def init():
    return (0, 0)


def quit():
    return None
