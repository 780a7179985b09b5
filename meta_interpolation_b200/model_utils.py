"""Model-plugin utilities kept for API compatibility (reference model_utils.py:272-305).

The B200 plugins do not peel parameter dictionaries level by level (that was
35 ms of pure Python per CAIN forward); ``extract_top_level_dict`` is
kept, with the reference's behaviour, for callers that still use it.
"""

_STRIPPED = ("layer_dict.", "block_dict.", "module-")


def extract_top_level_dict(current_dict):
    """Group ``{'a.b.c': t}`` by the first name level: ``{'a': {'b.c': t}}``; a key with a
    single level maps straight to its tensor.  Wrapper prefixes the reference strips
    (``layer_dict.``, ``block_dict.``, ``module-``) are removed first."""
    grouped = {}
    for full_name, tensor in current_dict.items():
        name = full_name
        for token in _STRIPPED:
            name = name.replace(token, "")
        head, _, rest = name.partition(".")
        if rest == "":
            if head not in grouped:
                grouped[head] = tensor
            else:
                grouped[head] = dict(grouped[head], **{"": tensor})
        else:
            bucket = grouped.get(head)
            grouped[head] = dict(bucket, **{rest: tensor}) if isinstance(bucket, dict) else {rest: tensor}
    return grouped
