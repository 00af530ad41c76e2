"""Console helpers with the reference's call signatures (acropolis/pprint.py:35-146).  ``print_error`` raises
instead of calling ``exit(1)`` (pprint.py:100) so that library users can catch it; the message format is kept."""
from sys import stdout, stderr

import acropolis_b200.flags as flags
from acropolis_b200.info import version, url

_max_verbose_level = 1
_use_color = True


class AcropolisError(SystemExit):
    """Raised by print_error; subclasses SystemExit so scripts that relied on exit(1) still terminate."""

    def __init__(self, msg):
        super().__init__(1)
        self.msg = msg

    def __str__(self):
        return self.msg


def _c(code):
    return code if _use_color else ""


def print_version():
    if flags.verbose:
        stdout.write(f"{_c(chr(27) + '[38;5;209m')}ACROPOLIS-B200 v{version} ({url}){_c(chr(27) + '[0m')}\n\n")


def print_Yf(Yf, header=("mean", "high", "low")):
    if not flags.verbose:
        print(*Yf.transpose().reshape(1, Yf.size)[0, :])
        return
    header = list(header) + [""] * (Yf.shape[1] - len(header))
    Yf = Yf.copy()
    Yf[Yf <= 1e-99] = 0
    labels = ['n', 'p', 'H2', 'H3', 'He3', 'He4', 'Li6', 'Li7', 'Be7']
    print(("\n{:^4}" + " | {:>12}" * Yf.shape[1]).format("", *header))
    print("-" * (5 + 15 * Yf.shape[1]))
    for j, label in enumerate(labels):
        line = "{:>4}".format(label) + "".join(" | {:11.6e}".format(v) for v in Yf[j])
        if label in ('n', 'H3', 'Be7'):
            line += "  [decayed]"
        print(line)


def print_error(error, loc="", eol="\n", flush=False):
    locf = f" ({loc})" if flags.debug and loc else ""
    stderr.write(f"{_c(chr(27) + '[1;31m')}ERROR  {_c(chr(27) + '[0m')}: {error}{locf}{eol}")
    if flush:
        stderr.flush()
    raise AcropolisError(error)


def print_warning(warning, loc="", eol="\n", flush=False):
    locf = f" ({loc})" if flags.debug and loc else ""
    stdout.write(f"{_c(chr(27) + '[1;33m')}WARNING{_c(chr(27) + '[0m')}: {warning}{locf}{eol}")
    if flush:
        stdout.flush()


def print_info(info, loc="", eol="\n", flush=False, verbose_level=None):
    global _max_verbose_level
    if verbose_level is None:
        verbose_level = _max_verbose_level
    _max_verbose_level = max(_max_verbose_level, verbose_level)
    locf = f" ({loc})" if flags.debug and loc else ""
    if flags.verbose and verbose_level >= _max_verbose_level:
        stdout.write(f"{_c(chr(27) + '[1;32m')}INFO   {_c(chr(27) + '[0m')}: {info}{locf}{eol}")
        if flush:
            stdout.flush()


def set_max_verbose_level(max_verbose_level=None):
    global _max_verbose_level
    _max_verbose_level = 1 if max_verbose_level is None else max_verbose_level


def disable_color():
    global _use_color
    _use_color = False
