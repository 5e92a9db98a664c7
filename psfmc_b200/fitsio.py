"""
Minimal FITS reader/writer (images + binary tables), numpy only.

The reference reads its inputs with ``astropy.io.fits.getdata/getheader``
(/root/reference/psfMC/utils.py:60-62,110-111) and writes its trace database
and posterior images through astropy (/root/reference/psfMC/database.py:42,
/root/reference/psfMC/analysis/images.py:97-100). astropy is not a dependency of
this package, so the subset of the FITS standard those call sites need is
implemented here: primary-HDU images (BITPIX 8/16/32/64/-32/-64, BSCALE/BZERO,
2880-byte blocks, optional gzip) and BINTABLE extensions with scalar and
fixed-length vector columns of the L/B/I/J/K/E/D/A types.
"""
import gzip
import io
from collections import OrderedDict

import numpy as np

BLOCK = 2880
CARD = 80

_BITPIX_DTYPE = {8: '>u1', 16: '>i2', 32: '>i4', 64: '>i8', -32: '>f4', -64: '>f8'}
_DTYPE_BITPIX = {'u1': 8, 'i2': 16, 'i4': 32, 'i8': 64, 'f4': -32, 'f8': -64}


class FITSFormatError(IOError):
    """Raised for anything that is not a readable FITS file. It is an IOError
    on purpose: the reference's mask loader tries FITS first and falls back to
    ds9 regions on IOError (/root/reference/psfMC/utils.py:87-90)."""


class Header(OrderedDict):
    """Ordered keyword -> value mapping; comments kept in ``.comments`` and
    COMMENT/HISTORY/blank cards kept (in order of appearance) in ``.commentary``."""

    def __init__(self, *args, **kwargs):
        super(Header, self).__init__(*args, **kwargs)
        self.comments = {}
        self.commentary = []

    def set(self, key, value=None, comment=None):
        self[key] = value
        if comment is not None:
            self.comments[key] = comment

    def copy(self):
        new = Header(self)
        new.comments = dict(self.comments)
        new.commentary = list(self.commentary)
        return new


def _open_bytes(source):
    if isinstance(source, (bytes, bytearray)):
        return bytes(source)
    try:
        with open(source, 'rb') as fobj:
            raw = fobj.read()
    except (OSError, TypeError) as err:
        raise FITSFormatError(str(err))
    if raw[:2] == b'\x1f\x8b':
        raw = gzip.decompress(raw)
    return raw


def _parse_value(text):
    text = text.strip()
    if not text:
        return None
    if text.startswith("'"):
        # string: ends at the first single quote not doubled
        out, i = [], 1
        while i < len(text):
            if text[i] == "'":
                if i + 1 < len(text) and text[i + 1] == "'":
                    out.append("'")
                    i += 2
                    continue
                break
            out.append(text[i])
            i += 1
        return ''.join(out).rstrip()
    token = text.split('/')[0].strip()
    if token == 'T':
        return True
    if token == 'F':
        return False
    try:
        return int(token)
    except ValueError:
        pass
    try:
        return float(token.replace('D', 'E').replace('d', 'e'))
    except ValueError:
        return token


def _split_comment(text):
    """Return the comment part of a value card body (after the value)."""
    in_str = False
    for i, ch in enumerate(text):
        if ch == "'":
            in_str = not in_str
        elif ch == '/' and not in_str:
            return text[i + 1:].strip()
    return None


def _parse_header(raw, offset):
    hdr = Header()
    pos = offset
    done = False
    while not done:
        block = raw[pos:pos + BLOCK]
        if len(block) < BLOCK:
            raise FITSFormatError('truncated FITS header')
        pos += BLOCK
        for i in range(0, BLOCK, CARD):
            card = block[i:i + CARD].decode('ascii', 'replace')
            key = card[:8].strip()
            if key == 'END':
                done = True
                break
            if key == 'HIERARCH' and '=' in card:
                long_key, _, rest = card[9:].partition('=')
                hdr[long_key.strip()] = _parse_value(rest)
            elif card[8:10] == '= ':
                hdr[key] = _parse_value(card[10:])
                comment = _split_comment(card[10:])
                if comment:
                    hdr.comments[key] = comment
            elif key in ('COMMENT', 'HISTORY', ''):
                if card.strip():
                    hdr.commentary.append((key, card[8:].rstrip()))
    return hdr, pos


def _data_size(hdr):
    naxis = hdr.get('NAXIS', 0)
    if naxis == 0:
        return 0
    count = 1
    for ax in range(1, naxis + 1):
        count *= hdr['NAXIS{:d}'.format(ax)]
    count = (count + hdr.get('PCOUNT', 0)) * hdr.get('GCOUNT', 1)
    return count * abs(hdr['BITPIX']) // 8


def _iter_hdus(raw):
    if raw[:6] != b'SIMPLE':
        raise FITSFormatError('not a FITS file (no SIMPLE card)')
    pos = 0
    while pos < len(raw):
        if not raw[pos:pos + 8].strip():
            break
        hdr, data_start = _parse_header(raw, pos)
        nbytes = _data_size(hdr)
        yield hdr, data_start, nbytes
        pos = data_start + ((nbytes + BLOCK - 1) // BLOCK) * BLOCK


def _image_from(raw, hdr, start, nbytes):
    naxis = hdr.get('NAXIS', 0)
    if naxis == 0:
        return None
    shape = tuple(hdr['NAXIS{:d}'.format(ax)] for ax in range(naxis, 0, -1))
    dtype = np.dtype(_BITPIX_DTYPE[hdr['BITPIX']])
    if start + nbytes > len(raw):
        raise FITSFormatError('truncated FITS data unit')
    data = np.frombuffer(raw, dtype=dtype, count=int(np.prod(shape)),
                         offset=start).reshape(shape)
    # native byte order, writable copy (astropy hands back big-endian views; the
    # values are identical and every consumer here only looks at values)
    data = data.astype(dtype.newbyteorder('='))
    bscale, bzero = hdr.get('BSCALE', 1), hdr.get('BZERO', 0)
    if bscale != 1 or bzero != 0:
        data = data * np.float64(bscale) + np.float64(bzero)
        if hdr['BITPIX'] in (8, 16):
            data = data.astype(np.float32)
    return data


def getheader(source, ext=0):
    """Header of HDU ``ext`` (cf. astropy.io.fits.getheader)."""
    raw = _open_bytes(source)
    for num, (hdr, _start, _nbytes) in enumerate(_iter_hdus(raw)):
        if num == ext:
            return hdr
    raise FITSFormatError('HDU {} not found'.format(ext))


def getdata(source, ext=None, header=False, **_ignored):
    """Image array of the first HDU that has data (or HDU ``ext``), like
    astropy.io.fits.getdata. Extra keyword arguments the reference passes
    (``ignore_missing_end``) are accepted and ignored."""
    raw = _open_bytes(source)
    for num, (hdr, start, nbytes) in enumerate(_iter_hdus(raw)):
        if ext is not None and num != ext:
            continue
        if hdr.get('XTENSION', 'IMAGE').strip() == 'BINTABLE':
            data = _table_from(raw, hdr, start)
        else:
            data = _image_from(raw, hdr, start, nbytes)
        if data is None and ext is None:
            continue
        return (data, hdr) if header else data
    raise FITSFormatError('no data found in FITS file')


# ---------------------------------------------------------------- writing --

def _format_card(key, value, comment=None):
    if len(str(key)) > 8:
        # keywords longer than eight characters go into a HIERARCH card, like astropy
        # writes them (e.g. the reference's '2SER_index' posterior summaries)
        text = str(value).replace("'", "''") if not isinstance(
            value, (bool, np.bool_, int, np.integer, float, np.floating)) else None
        body = "'{}'".format(text) if text is not None else (
            ('T' if value else 'F') if isinstance(value, (bool, np.bool_)) else repr(value))
        card = 'HIERARCH {} = {}'.format(key, body)
        if comment and len(card) + 3 + len(str(comment)) <= CARD:
            card += ' / ' + str(comment)
        return card[:CARD].ljust(CARD)
    key = str(key).upper()[:8]
    if isinstance(value, (bool, np.bool_)):
        body = '{:>20s}'.format('T' if value else 'F')
    elif isinstance(value, (int, np.integer)):
        body = '{:>20d}'.format(int(value))
    elif isinstance(value, (float, np.floating)):
        text = repr(float(value)).upper()
        if 'E' not in text and '.' not in text and 'N' not in text:
            text += '.'
        body = '{:>20s}'.format(text)
    elif value is None:
        body = ''
    else:
        text = str(value).replace("'", "''")
        body = "'{:<8s}'".format(text[:67])
    card = '{:<8s}= {}'.format(key, body)
    if comment:
        card += ' / ' + str(comment)
    return card[:CARD].ljust(CARD)


def _header_bytes(cards):
    text = ''.join(cards) + 'END'.ljust(CARD)
    pad = (-len(text)) % BLOCK
    return (text + ' ' * pad).encode('ascii', 'replace')


def _pad_block(data_bytes, fill=b'\x00'):
    pad = (-len(data_bytes)) % BLOCK
    return data_bytes + fill * pad


_STRUCTURAL = ('SIMPLE', 'BITPIX', 'NAXIS', 'EXTEND', 'XTENSION', 'PCOUNT',
               'GCOUNT', 'TFIELDS', 'BSCALE', 'BZERO', 'END')


def _user_cards(header):
    cards = []
    if header is None:
        return cards
    comments = getattr(header, 'comments', {})
    for key, value in header.items():
        ukey = str(key).upper()
        if ukey in _STRUCTURAL or ukey.startswith('NAXIS') or \
                ukey[:5] in ('TTYPE', 'TFORM', 'TUNIT', 'TDIM'):
            continue
        comment = comments.get(key)
        if isinstance(value, tuple):
            value, comment = value
        cards.append(_format_card(key, value, comment))
    for key, text in getattr(header, 'commentary', []):
        cards.append('{:<8s}{}'.format(key, text)[:CARD].ljust(CARD))
    return cards


def image_hdu_bytes(data, header=None, primary=True):
    data = np.asarray(data)
    if data.dtype == np.bool_:
        data = data.astype(np.uint8)
    code = data.dtype.str[1:]
    if code not in _DTYPE_BITPIX:
        data = data.astype(np.float64)
        code = 'f8'
    bitpix = _DTYPE_BITPIX[code]
    cards = [_format_card('SIMPLE', True, 'conforms to FITS standard')
             if primary else _format_card('XTENSION', 'IMAGE')]
    cards.append(_format_card('BITPIX', bitpix))
    cards.append(_format_card('NAXIS', data.ndim))
    for ax in range(data.ndim):
        cards.append(_format_card('NAXIS{:d}'.format(ax + 1),
                                  data.shape[data.ndim - 1 - ax]))
    if primary:
        cards.append(_format_card('EXTEND', True))
    else:
        cards += [_format_card('PCOUNT', 0), _format_card('GCOUNT', 1)]
    cards += _user_cards(header)
    payload = data.astype(data.dtype.newbyteorder('>')).tobytes()
    return _header_bytes(cards) + _pad_block(payload)


def writeto(filename, data, header=None, overwrite=True, **_ignored):
    """Write a single-image FITS file (cf. astropy.io.fits.writeto)."""
    mode = 'wb' if overwrite else 'xb'
    with open(filename, mode) as fobj:
        fobj.write(image_hdu_bytes(data, header, primary=True))


# ------------------------------------------------------------ bin tables --

_TFORM_NP = {'L': 'u1', 'B': 'u1', 'I': '>i2', 'J': '>i4', 'K': '>i8',
             'E': '>f4', 'D': '>f8'}


def _tform_for(arr):
    kind = arr.dtype.kind
    repeat = int(np.prod(arr.shape[1:])) if arr.ndim > 1 else 1
    if kind == 'b':
        return repeat, 'L', 'u1'
    if kind in 'iu':
        if arr.dtype.itemsize <= 2 and kind == 'i':
            return repeat, 'I', '>i2'
        if arr.dtype.itemsize <= 4 and kind == 'i':
            return repeat, 'J', '>i4'
        return repeat, 'K', '>i8'
    if kind == 'f':
        if arr.dtype.itemsize == 4:
            return repeat, 'E', '>f4'
        return repeat, 'D', '>f8'
    if kind in 'SU':
        width = max(1, max(len(str(s)) for s in arr.ravel()) if arr.size else 1)
        return width, 'A', 'S{:d}'.format(width)
    raise TypeError('unsupported column dtype {}'.format(arr.dtype))


def write_table(filename, columns, header=None, overwrite=True):
    """
    Write an empty primary HDU followed by one BINTABLE extension.

    :param columns: ordered mapping name -> 1-D (or (nrows, k)) array
    :param header: mapping of extra header keywords for the table HDU; values
        may be ``(value, comment)`` tuples as produced by the reference's
        ``annotate_metadata`` (/root/reference/psfMC/database.py:90-109)
    """
    names = list(columns.keys())
    arrays = [np.asarray(columns[name]) for name in names]
    nrows = arrays[0].shape[0] if arrays else 0
    fields, cards = [], []
    for arr in arrays:
        if arr.shape[0] != nrows:
            raise ValueError('all table columns must have the same length')
    row_dtype = []
    for num, (name, arr) in enumerate(zip(names, arrays)):
        repeat, code, npcode = _tform_for(arr)
        fields.append((repeat, code, npcode))
        if code == 'A':
            row_dtype.append(('f{:d}'.format(num), npcode))
        elif repeat == 1:
            row_dtype.append(('f{:d}'.format(num), npcode))
        else:
            row_dtype.append(('f{:d}'.format(num), npcode, (repeat,)))
    rec = np.zeros(nrows, dtype=np.dtype(row_dtype))
    for num, arr in enumerate(arrays):
        repeat, code, npcode = fields[num]
        col = arr
        if code == 'L':
            col = np.where(arr, ord('T'), ord('F')).astype('u1')
        elif code == 'A':
            col = np.char.encode(arr.astype(str), 'ascii') if arr.dtype.kind == 'U' else arr
        if repeat > 1 and code != 'A':
            col = col.reshape(nrows, repeat)
        rec['f{:d}'.format(num)] = col
    cards.append(_format_card('XTENSION', 'BINTABLE', 'binary table extension'))
    cards.append(_format_card('BITPIX', 8))
    cards.append(_format_card('NAXIS', 2))
    cards.append(_format_card('NAXIS1', rec.dtype.itemsize))
    cards.append(_format_card('NAXIS2', nrows))
    cards.append(_format_card('PCOUNT', 0))
    cards.append(_format_card('GCOUNT', 1))
    cards.append(_format_card('TFIELDS', len(names)))
    for num, name in enumerate(names):
        repeat, code, _ = fields[num]
        cards.append(_format_card('TTYPE{:d}'.format(num + 1), name))
        tform = '{:d}{}'.format(repeat, code) if (repeat != 1 or code == 'A') else code
        cards.append(_format_card('TFORM{:d}'.format(num + 1), tform))
    cards += _user_cards(header)
    primary = _header_bytes([
        _format_card('SIMPLE', True, 'conforms to FITS standard'),
        _format_card('BITPIX', 8), _format_card('NAXIS', 0),
        _format_card('EXTEND', True)])
    mode = 'wb' if overwrite else 'xb'
    with open(filename, mode) as fobj:
        fobj.write(primary)
        fobj.write(_header_bytes(cards))
        fobj.write(_pad_block(rec.tobytes()))


class Table(OrderedDict):
    """name -> column array, plus ``.meta`` (the table HDU's extra keywords)."""

    def __init__(self, *args, **kwargs):
        super(Table, self).__init__(*args, **kwargs)
        self.meta = Header()

    @property
    def colnames(self):
        return list(self.keys())

    def __len__(self):
        for col in self.values():
            return len(col)
        return 0

    def select(self, row_mask):
        out = Table((name, col[row_mask]) for name, col in self.items())
        out.meta = self.meta.copy()
        return out


def _table_from(raw, hdr, start):
    nrows, rowlen = hdr['NAXIS2'], hdr['NAXIS1']
    fields = []
    for num in range(1, hdr['TFIELDS'] + 1):
        tform = str(hdr['TFORM{:d}'.format(num)]).strip()
        digits = ''.join(ch for ch in tform if ch.isdigit())
        code = tform[len(digits)]
        repeat = int(digits) if digits else 1
        name = str(hdr.get('TTYPE{:d}'.format(num), 'col{:d}'.format(num))).strip()
        if code == 'A':
            fields.append((name, 'S{:d}'.format(repeat)))
        elif repeat == 1:
            fields.append((name, _TFORM_NP[code]))
        else:
            fields.append((name, _TFORM_NP[code], (repeat,)))
        fields[-1] = fields[-1] + ((code,),)
    dtype = np.dtype([f[:-1] for f in fields])
    if dtype.itemsize != rowlen:
        raise FITSFormatError('BINTABLE row length mismatch')
    rec = np.frombuffer(raw, dtype=dtype, count=nrows, offset=start)
    table = Table()
    for field in fields:
        name, code = field[0], field[-1][0]
        col = rec[name]
        if code == 'L':
            col = col == ord('T')
        elif code == 'A':
            col = np.char.decode(col, 'ascii')
        else:
            col = col.astype(col.dtype.newbyteorder('='))
        table[name] = col
    for key, value in hdr.items():
        if key in _STRUCTURAL or key.startswith('NAXIS') or \
                key[:5] in ('TTYPE', 'TFORM', 'TUNIT', 'TDIM'):
            continue
        table.meta[key] = value
        if key in hdr.comments:
            table.meta.comments[key] = hdr.comments[key]
    return table


def read_table(filename, ext=1):
    """Read a BINTABLE extension into a :class:`Table`."""
    return getdata(filename, ext=ext)
