"""Stand-in for ``typing_validation`` (test infrastructure; see hashcons.py here)."""


def validate(val, t=None):
    return True


def can_validate(t):
    return True
