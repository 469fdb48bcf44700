"""Host-side mirror of the reference's search request (www/FlaskApp/FlaskApp/patmatch.py).

Same function names, argument meaning and result shapes as the reference module for
the search path, with the two shell-outs replaced:

  patmatch_to_nrgrep.pl  (patmatch.py:291-297)  -> patmatchdocker_b200.pattern.convert
  nrgrep_coords          (patmatch.py:733-743)  -> Engine.search on the HBM-resident dataset
  generate_sequence_index.pl (patmatch.py:202)  -> get_record_offset below (numpy)

Out of scope (not on the search path): S3 upload / download URLs, temp-file cleanup,
Flask routing, get_config / get_sequence.
"""
import re
import threading

import numpy as np

from . import pattern as _pattern
from ._native import Engine, NativeError

MIN_TOKEN = 3
MINHITS = 500
MAXHITS = 100000
DEFAULT_MAXHITS = 500


# --------------------------------------------------------------------------- request checks
def check_pattern(pattern, seqtype):
    """patmatch.py:241-267."""
    if seqtype in ['pep', 'protein']:
        if 'u' in pattern.lower():
            return 'Invalid peptide character found in pattern.'
    else:
        if any(x in pattern.upper() for x in ('E', 'F', 'I', 'J', 'L', 'O', 'P', 'Q', 'Z')):
            return 'Invalid nucleotide character found in pattern.'
    tokens, counting = 0, True
    for x in pattern:
        if x in '([{':
            if counting:
                tokens += 1
            counting = False
        elif x in ')]}':
            counting = True
        elif counting:
            tokens += 1
    if '{' in pattern:
        return ''
    if tokens < MIN_TOKEN:
        return "Your pattern is shorter than the minimum number of " + str(MIN_TOKEN) + " residues."
    return ''


def cleanup_pattern(pattern):
    """patmatch.py:363-371."""
    for a, b in (('%28', '('), ('%29', ')'), ('%7B', '{'), ('%7D', '}'), ('%5B', '['), ('%5D', ']'),
                 ('%2C', ','), ('%5E', '^')):
        pattern = pattern.replace(a, b)
    return pattern


def process_pattern(pattern, seqtype, strand, insertion, deletion, substitution, mismatch):
    """patmatch.py:270-316 -> (nrgrep pattern, reverse-complement pattern or '', '-k' argument)."""
    if seqtype is None:
        seqtype = 'pep'
    if seqtype in ['pep', 'protein']:
        option = _pattern.PEPTIDE
    elif strand and 'complement' in strand.lower():
        option = _pattern.COMPLEMENT
    else:
        option = _pattern.NUCLEOTIDE
    converted = _pattern.convert(pattern, option)
    comp_pattern = ""
    if seqtype.lower() in ['dna', 'nuc'] and (strand is None or strand.startswith('Both')):
        comp_pattern = _pattern.convert(converted, _pattern.COMPLEMENT)
    mismatch_option = ""
    if insertion and insertion.startswith('insertion'):
        mismatch_option += 'i'
    if deletion and deletion.startswith('deletion'):
        mismatch_option += 'd'
    if substitution and substitution.startswith('substitution'):
        mismatch_option += 's'
    if mismatch_option == '':
        mismatch_option = 'ids'
    if mismatch is None:
        mismatch = 0
    return converted, comp_pattern, str(mismatch) + mismatch_option


# --------------------------------------------------------------------------- sequence index
def get_record_offset(data):
    """generate_sequence_index.pl via patmatch.py:197-215.

    data: the .seq file bytes.  For every header line '>name ...' two entries: the byte
    offset of the header (named '>name') and the offset of the sequence line (named 'name').
    """
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    n = buf.size
    nl = np.flatnonzero(buf == 10)
    starts = np.concatenate(([0], nl + 1))
    starts = starts[starts < n]
    heads = starts[buf[starts] == ord('>')]
    offsets, names = [], {}
    raw = buf.tobytes() if n < (1 << 31) else None
    for h in heads.tolist():
        e = int(nl[np.searchsorted(nl, h)]) + 1 if np.searchsorted(nl, h) < nl.size else n   # line incl. '\n'
        line = (raw[h:e] if raw is not None else buf[h:e].tobytes()).decode('latin-1')
        m = re.match(r'^>(\S+)', line)
        if not m:
            continue
        for off, nm in ((h, '>' + m.group(1)), (h + len(line), m.group(1))):
            offsets.append(off)
            names[off] = nm
    return offsets, names


def get_name_offset(offSet, recordOffSetList):
    """patmatch.py:218-238 (kept verbatim in behaviour, including its binary-search quirks)."""
    low, high = 0, len(recordOffSetList) - 1
    while high > low:
        middle = int((low + high) / 2)
        if recordOffSetList[middle] == offSet:
            return offSet
        elif high - low == 1:
            return recordOffSetList[high] if offSet >= recordOffSetList[high] else recordOffSetList[low]
        elif recordOffSetList[middle] < offSet:
            low = middle
        elif recordOffSetList[middle] > offSet:
            high = middle - 1
    return recordOffSetList[low]


def set_seq_length(data_text):
    """patmatch.py:374-400 -> {name: biological length}."""
    lengths, seq, name = {}, '', ''
    for line in data_text.splitlines(True):
        if line.startswith('>'):
            if name != '':
                canon = name.rstrip(',')
                lengths[canon] = len(seq[:-1] if seq.endswith('*') else seq)
            name = line.replace('>', '').split(' ')[0].rstrip(',')
            seq = ''
        else:
            seq += line.strip()
    if name and seq:
        canon = name.rstrip(',')
        lengths[canon] = len(seq[:-1] if seq.endswith('*') else seq)
    return lengths


def find_exclusion_offset(pattern):
    """patmatch.py:403-446."""
    tokens = re.findall(r'\[[^\]]+\]|.(?:[*+?]|\{\d*(?:,\d*)?\})?', pattern)
    excl_idx = next((i for i, t in enumerate(tokens) if t.startswith('[^')), None)
    if excl_idx is None:
        return None
    offset = 0
    for tok in tokens[:excl_idx]:
        if tok.startswith('['):
            offset += 1
        elif len(tok) > 1:
            quant = tok[1:]
            if quant in ('*', '?'):
                reps = 0
            elif quant == '+':
                reps = 1
            elif quant.startswith('{'):
                q = quant.strip('{}').split(',')
                try:
                    reps = int(q[0]) if q[0] else 0
                except (ValueError, IndexError):
                    reps = 0
            else:
                reps = 1
            offset += reps
    return offset


# --------------------------------------------------------------------------- engine output
def format_output(hits, data, banner):
    """The stdout of one nrgrep_coords run (recSearchFile @402250): banner, '[b, e]: text', total."""
    raw = data if isinstance(data, (bytes, bytearray)) else data.tobytes()
    lines = [banner]
    for b, e in hits:
        lines.append("[%d, %d]: %s" % (b, e, raw[b:e].decode('latin-1')))
    lines.append("Total: %d matching records" % len(hits))
    return "\n".join(lines) + "\n"


def process_output(recordOffSetList, seqNm4offSet, output, datafile, maxhits, begMatch, endMatch,
                   original_pattern, data_text=None, locus=None):
    """patmatch.py:449-674 -> (rows, uniqueHits, totalHits, file_lines).

    `output` is nrgrep_coords stdout text (format_output above); rows are the dicts the
    reference returns under "hits"; file_lines is the TSV the reference writes for download.
    """
    seqNm2length = set_seq_length(data_text) if endMatch == 1 and data_text is not None else {}
    exclusion_positions = [(find_exclusion_offset(m.string[:m.start()]), set(m.group(1)))
                           for m in re.finditer(r'\[\^([^\]]+)\]', original_pattern)]
    name2data = locus or {}
    seqNm2chr, seqNm2orfs = {}, {}
    if 'Not' in datafile and data_text is not None:
        for line in data_text.splitlines():
            if line.startswith('>'):
                pieces = line.strip().replace('>', '').split(' ')
                seqName = pieces[0].replace(',', '')
                seqNm2chr[seqName] = pieces[2]
                seqNm2orfs[seqName] = line.strip().split('between ')[1].replace('and', '-')
    if maxhits is None:
        maxhits = DEFAULT_MAXHITS
    elif str(maxhits).isdigit():
        maxhits = int(maxhits)
    elif str(maxhits).lower() in ['no limit', 'no+limit']:
        maxhits = MAXHITS
    else:
        maxhits = DEFAULT_MAXHITS

    data, totalHits, uniqueHits, hitCount4seqNm = [], 0, 0, {}
    for line in output.split('\n'):
        if not line.startswith('['):
            continue
        line = line.replace('[', '').replace(']', '').replace(':', '').replace(',', '')
        pieces = line.split(' ')
        if len(pieces) < 3:
            continue
        beg, end, matchingPattern = int(pieces[0]), int(pieces[1]), pieces[2]
        if any(pos is not None and pos < len(matchingPattern) and matchingPattern[pos] in chars
               for pos, chars in exclusion_positions):
            continue
        offSet = get_name_offset(beg, recordOffSetList)
        seqBeg, seqEnd = beg - offSet + 1, end - offSet
        seqNm = seqNm4offSet.get(offSet, None)
        if seqNm is None:
            continue
        if begMatch == 1 and seqBeg != 1:
            continue
        if endMatch == 1:
            length = seqNm2length.get(seqNm)
            if length is None or seqEnd != length:
                continue
        if seqNm.startswith('>'):
            continue
        if seqNm.endswith(','):
            seqNm = seqNm.rstrip(seqNm[-1])
        if 'Not' in datafile:
            pieces = seqNm.split(':')
            if len(pieces) < 2:
                continue
            num = int(pieces[1].split('-')[0])
            seqBeg, seqEnd = seqBeg + num - 1, seqEnd + num - 1
            if seqNm not in seqNm2chr or seqNm not in seqNm2orfs:
                continue
            row = "\t".join((str(seqNm2orfs.get(seqNm)), str(seqBeg), str(seqEnd), matchingPattern,
                             str(seqNm2chr.get(seqNm)), seqNm))
        else:
            gene, sgdid, desc = name2data.get(seqNm, ('', '', ''))
            row = "\t".join((seqNm, str(seqBeg), str(seqEnd), matchingPattern, gene, sgdid, desc))
        if seqNm not in hitCount4seqNm:
            uniqueHits += 1
        if totalHits >= maxhits:
            break
        hitCount4seqNm[seqNm] = hitCount4seqNm.get(seqNm, 0) + 1
        totalHits += 1
        data.append(row)

    if 'Not' in datafile:
        file_lines = ["Chromosome\tBetweenORFtoORF\tHitNumber\tMatchPattern\tMatchStartCoord\tMatchStopCoord\n"]
    elif 'orf_' in datafile:
        file_lines = ["Feature Name\tGene Name\tHitNumber\tMatchPattern\tMatchStartCoord\tMatchStopCoord\tLocusInfo\n"]
    else:
        file_lines = ["Sequence Name\tHitNumber\tMatchPattern\tMatchStartCoord\tMatchStopCoord\n"]
    newData = []
    data.sort()
    for row in data:
        if 'Not' in datafile:
            orfs, beg, end, matchPattern, chrom, seqNm = row.split('\t')
            count = hitCount4seqNm[seqNm]
            orfs = orfs.strip()
            newData.append({'orfs': orfs, 'chr': chrom, 'beg': beg, 'end': end, 'count': count,
                            'seqname': seqNm, 'matchingPattern': matchPattern})
            # (the reference builds this line but only appends lines in its else-branch)
        else:
            seqNm, beg, end, matchPattern, gene, sgdid, desc = row.split('\t')
            count = hitCount4seqNm.get(seqNm, 0)
            if sgdid != "":
                if gene == seqNm:
                    gene = ""
                newData.append({'seqname': seqNm, 'beg': beg, 'end': end, 'count': count,
                                'matchingPattern': matchPattern, 'gene_name': gene, 'sgdid': sgdid, 'desc': desc})
                line = "\t".join((seqNm, gene, str(count), matchPattern, beg, end, desc)) + "\n"
            else:
                newData.append({'seqname': seqNm, 'gene_name': gene, 'sgdid': sgdid, 'beg': beg, 'end': end,
                                'count': count, 'matchingPattern': matchPattern, 'desc': desc})
                line = "\t".join((seqNm, str(count), matchPattern, beg, end)) + "\n"
            file_lines.append(line)
    return newData, uniqueHits, totalHits, file_lines


# --------------------------------------------------------------------------- the request
class PatMatch:
    """run_patmatch (patmatch.py:677-765) against datasets resident on one GPU."""

    def __init__(self, device=0, engine=None):
        self.engine = engine or Engine(device)
        self.datasets = {}
        self._lock = threading.RLock()       # request threads of one WSGI process share the engine and the dataset table

    def add_dataset(self, name, data, locus=None):
        """name as the reference's '<dataset>.seq' file name; data = file bytes."""
        raw = bytes(data) if not isinstance(data, bytes) else data
        offsets, names = get_record_offset(raw)
        with self._lock:
            self.datasets[name] = {"ds": self.engine.load_dataset(raw), "raw": raw, "offsets": offsets,
                                   "names": names, "locus": locus, "text": None}

    def _nrgrep_request(self, entry, patterns, option):
        """Both nrgrep_coords runs of a request (patmatch.py:733-735 and :739-743) in ONE pass over the dataset
        (pm_search_request): the stdout texts the reference would have read, in the order it concatenates them."""
        banner = "SIMPLE search" if int(re.match(r"\d+", option).group()) == 0 else "ESIMPLE search"
        lists = self.engine.search_request(entry["ds"], patterns, option)
        return [format_output([(int(b), int(e)) for b, e in hits], entry["raw"], banner) for hits in lists]

    def _nrgrep(self, entry, pattern, option):
        """One nrgrep_coords run: (stdout text the reference would have read, hit array)."""
        hits = self.engine.search(entry["ds"], pattern, option)
        banner = "SIMPLE search" if int(re.match(r"\d+", option).group()) == 0 else "ESIMPLE search"
        return format_output([(int(b), int(e)) for b, e in hits], entry["raw"], banner), hits

    def run_patmatch(self, pattern, seqtype=None, dataset=None, strand=None, insertion=None, deletion=None,
                     substitution=None, mismatch=None, max_hits=None):
        seqtype_eff = seqtype if seqtype is not None else 'pep'
        if dataset:
            dataset = dataset + ".seq"
        else:
            dataset = "orf_dna.seq" if seqtype_eff in ['dna', 'nuc'] else "orf_pep.seq"
        if dataset not in self.datasets:
            raise KeyError("dataset %r is not loaded" % dataset)
        entry = self.datasets[dataset]
        pattern = cleanup_pattern(pattern)
        begMatch = endMatch = 0
        if pattern.startswith('<'):
            begMatch, pattern = 1, pattern.replace('<', '')
        elif pattern.endswith('>'):
            endMatch, pattern = 1, pattern.replace('>', '')
        error = check_pattern(pattern, seqtype_eff)
        if error:
            return {"error": error}
        conv, comp, option = process_pattern(pattern, seqtype, strand, insertion, deletion, substitution, mismatch)
        try:
            with self._lock:
                outs = self._nrgrep_request(entry, [conv, comp] if comp else [conv], option)
            output = outs[0] if len(outs) == 1 else outs[0] + "\n" + outs[1]
        except NativeError as e:
            return {"error": str(e)}
        if (endMatch == 1 or 'Not' in dataset) and entry["text"] is None:
            entry["text"] = entry["raw"].decode('latin-1')
        rows, uniqueHits, totalHits, file_lines = process_output(
            entry["offsets"], entry["names"], output, dataset, max_hits, begMatch, endMatch, conv,
            data_text=entry["text"], locus=entry["locus"])
        return {"hits": rows, "uniqueHits": uniqueHits, "totalHits": totalHits, "downloadUrl": "",
                "error_message": "", "download_lines": file_lines}
