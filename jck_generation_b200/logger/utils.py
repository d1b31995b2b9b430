def time_to_str(time_diff: float) -> str:
    hours, rest = divmod(time_diff, 3600)
    minutes, seconds = divmod(rest, 60)
    return f'{hours}h {minutes}m {seconds}'
