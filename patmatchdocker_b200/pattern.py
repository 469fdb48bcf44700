"""PatMatch pattern -> nrgrep pattern conversion (host side of the hot path).

Restates www/bin/patmatch_to_nrgrep.pl of the reference (called by
www/FlaskApp/FlaskApp/patmatch.py:291-297 through os.popen) so that the search
request never leaves the process.  Behaviour is pinned against the Perl script
by tests/golden/pattern_golden.json (tests/golden/make_golden.py).

Classes (patmatch_to_nrgrep.pl:11-14):  "-n" nucleotide, "-p" peptide,
"-c" reverse-complemented nucleotide.
"""
import re

NUCLEOTIDE, PEPTIDE, COMPLEMENT = "-n", "-p", "-c"

# sub_characters (patmatch_to_nrgrep.pl:515-542)
_PEPTIDE_CLASSES = (("J", "[IFVLWMAGCY]"), ("O", "[TSHEDQNKR]"), ("B", "[DN]"), ("Z", "[EQ]"))
_NUCLEOTIDE_CLASSES = (("R", "[AG]"), ("Y", "[CT]"), ("S", "[GC]"), ("W", "[AT]"), ("M", "[AC]"),
                       ("K", "[GT]"), ("V", "[ACG]"), ("H", "[ACT]"), ("D", "[AGT]"), ("B", "[CGT]"))
# complement_nucleotides (patmatch_to_nrgrep.pl:723)
_COMPLEMENT = str.maketrans("ATCGRYSWMKVHDB", "TAGCYRSWKMBDHV")
_OPENER = {")": "(", "]": "[", "}": "{"}


class PatternError(ValueError):
    pass


def convert(pattern, cls):
    """patmatch_to_nrgrep.pl <cls> '<pattern>' -> the text it prints (process_pattern, :90-98)."""
    if cls not in (NUCLEOTIDE, PEPTIDE, COMPLEMENT):
        raise PatternError("Invalid class.")          # check_class, :51-70
    p = re.sub(r"\s", "", pattern).upper()            # prepare_pattern, :119-129
    if cls == COMPLEMENT:
        p = reverse_complement(p)
    p = p.replace("X", ".") if cls == PEPTIDE else p.replace("N", ".").replace("X", ".")   # fix_wildcards
    p = _expand_repetitions(p)
    for letter, group in (_PEPTIDE_CLASSES if cls == PEPTIDE else _NUCLEOTIDE_CLASSES):
        p = p.replace(letter, group)
    p = _flatten_brackets(p)
    return _finalize(p)


def _expand_repetitions(p):
    """fix_repetitions (:184-211): X{m}, X{m,}, X{,n}, X{m,n} -> XX..X?X? / X*"""
    if "{" not in p:
        return p
    out = []                                           # list of pattern pieces, like @nrgrep
    for ch in p:
        if ch != "}":
            out.append(ch)
            continue
        # extract_repitition_information (:261-281): pop back to '{'
        info = []
        while True:
            if not out:
                raise PatternError("unbalanced '{'")
            c = out.pop()
            if c == "{":
                break
            info.insert(0, c)
        info = "".join(info)
        unit = _pop_repeat_unit(out)
        lower, upper = _repeat_bounds(info)
        if upper == -1:
            out.append(unit * lower + unit + "*")
        else:
            out.append(unit * lower + (unit + "?") * max(upper - lower, 0))
    return "".join(out)


def _pop_repeat_unit(out):
    """extract_repeat_pattern (:307-353): the atom or bracketed group a repetition applies to."""
    if not out:
        raise PatternError("repetition without operand")
    ch = out.pop()
    if ch not in (")", "]"):
        return ch
    closer, opener = ch, _OPENER[ch]
    depth, unit = 1, [ch]
    while depth > 0:
        if not out:
            raise PatternError("unbalanced group")
        ch = out.pop()
        unit.insert(0, ch)
        if ch == closer:
            depth += 1
        elif ch == opener:
            depth -= 1
    return "".join(unit)


def _num(s):
    """Perl numification of a string: leading digits, else 0."""
    m = re.match(r"\s*(\d+)", s or "")
    return int(m.group(1)) if m else 0


def _repeat_bounds(info):
    """process_repeat_info (:410-438); -1 means unbounded."""
    parts = info.split(",")
    while parts and parts[-1] == "":                   # Perl's split drops trailing empty fields
        parts.pop()
    parts += ["", ""]
    if re.match(r"^,\d+", info):
        return 0, _num(parts[1])
    if re.search(r"\d+,$", info):
        return _num(parts[0]), -1
    if re.match(r"^\d+$", info):
        return _num(info), _num(info)
    if re.match(r"^\d+,\d+$", info):
        return _num(parts[0]), _num(parts[1])
    return 0, 0


def _flatten_brackets(p):
    """remove_nested_brackets (:570-625): one bracket level, no repeated member."""
    out, depth, seen = [], 0, set()
    for ch in p:
        if ch == "[":
            if depth == 0:
                out.append(ch)
            depth += 1
        elif ch == "]":
            depth = max(depth - 1, 0)
            if depth == 0:
                out.append(ch)
                seen = set()
        elif depth == 0:
            out.append(ch)
        elif ch not in seen:
            out.append(ch)
            seen.add(ch)
    return "".join(out)


def _finalize(p):
    """finalize_pattern (:646-671): anchors outside, everything else in parentheses."""
    head, tail = p.startswith("<"), p.endswith(">")
    if head and tail:
        return "^(" + p.replace("<", "", 1).replace(">", "", 1) + ")$"
    if head:
        return "^(" + p.replace("<", "", 1) + ")"
    if tail:
        return "(" + p.replace(">", "", 1) + ")$"
    return "(" + p + ")"


def reverse_complement(p):
    """get_reverse_complement (:691-699) = complement_nucleotides + reverse_pattern."""
    p = p.translate(_COMPLEMENT)
    if p.startswith("<"):
        p = ">" + p[1:]
    if p.endswith(">"):
        p = p[:-1] + "<"
    chars, out = list(p), []
    while chars:
        ch = chars.pop()
        out.append(_pop_group(ch, chars) if ch in _OPENER else ch)
    return "".join(out)


def _pop_group(closer, chars):
    """extract_group (:806-866): reverse one (), [] or X{..} group, recursively."""
    opener = _OPENER[closer]
    group, inner = [closer], []
    while True:
        if not chars:
            raise PatternError("unbalanced group")
        ch = chars.pop()
        if ch == opener:
            if opener != "{":
                group.insert(0, "".join(inner))
                group.insert(0, ch)
            else:
                group.insert(0, ch)
                if not chars:
                    raise PatternError("repetition without operand")
                rep = chars.pop()
                group.insert(0, _pop_group(rep, chars) if rep in ("]", ")") else rep)
            break
        if ch in _OPENER:
            inner.append(_pop_group(ch, chars))
        elif closer == "}":
            group.insert(0, ch)
        else:
            inner.append(ch)
    return "".join(group)
