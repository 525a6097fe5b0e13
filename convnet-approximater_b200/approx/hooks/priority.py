"""Hook priorities: the mmcv-style level names the reference's configs may use (src/approx/hooks/priority.py)."""
from enum import Enum

Priority = Enum("Priority", dict(HIGHEST=0, VERY_HIGH=10, HIGH=30, ABOVE_NORMAL=40, NORMAL=50, BELOW_NORMAL=60, LOW=70,
                                 VERY_LOW=90, LOWEST=100))


def get_priority(priority) -> int:
    """int in [0, 100], a level name (any case) or a Priority member -> its integer value."""
    if isinstance(priority, Priority):
        return priority.value
    if isinstance(priority, bool):
        raise TypeError('priority must be an integer or Priority enum value')
    if isinstance(priority, int):
        if not 0 <= priority <= 100:
            raise ValueError('priority must be between 0 and 100')
        return priority
    if isinstance(priority, str):
        return Priority[priority.upper()].value
    raise TypeError('priority must be an integer or Priority enum value')
