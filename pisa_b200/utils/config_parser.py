"""Pipeline ``.cfg`` parsing, so that the reference's existing config files select the B200 services.

Restates the subset of pisa/utils/config_parser.py (reference :1-230 docstring, :303-958,
:1156-1500) that pipeline configs of the hot path use:

  * ``#include <resource> [as <section>]`` inlining (:1213-1214,1324-1500),
  * ``${section:key}`` extended interpolation, case-sensitive option names (:1255-1258),
  * ``[binning]`` entries ``<name>.order`` + one ``OneDimBinning`` kwargs dict per dimension (:732-748),
  * ``[pipeline]``: ``order``, ``name``, ``param_selections``, ``output_binning``, ``output_key`` (:752-795),
  * one ``[stage.service]`` section each: ``calc_mode`` / ``apply_mode`` (binning names are replaced by
    the parsed binning), ``*_names`` lists, ``param.[selector.]name = value [+/- sigma] [* units.X]``
    with ``.fixed`` / ``.range`` / ``.prior`` attributes (:845-905, parse_param :446-560); a param
    that already exists in an earlier stage is linked, not re-created (:862-887).

Priors are recorded as plain descriptions (kind / gaussian sigma) -- minimisation is out of scope.
"""
import re
from collections import OrderedDict
from configparser import ExtendedInterpolation, NoSectionError, RawConfigParser
from io import StringIO

import numpy as np

from pisa_b200 import FTYPE
from pisa_b200.utils.resources import find_resource
from pisa_b200.utils.units import Quantity, parse_quantity, ureg

__all__ = ["PISAConfigParser", "parse_pipeline_config", "parse_param", "parse_string_literal", "split",
           "PARAM_RE", "PARAM_ATTRS", "STAGE_SEP"]

PARAM_RE = re.compile(r"^param\.(?P<subfields>(([^.\s]+)(\.|$))+)", re.IGNORECASE)
PARAM_ATTRS = ["range", "prior", "fixed", "tex", "scales_as_log"]
STAGE_SEP = "."
INCLUDE_RE = re.compile(r"^\s*#include\s+(?P<include>\S.*?)\s*$")
INCLUDE_AS_RE = re.compile(r"^(?P<file>.+?)\s+as\s+(?P<as>\S+)$")

_EVAL_NS = dict(np=np, numpy=np, inf=np.inf, units=ureg, ureg=ureg, FTYPE=FTYPE)


def split(string, sep=","):
    return [x.strip() for x in str(string).split(sep) if x.strip()]


def parse_string_literal(string):
    low = string.strip().lower()
    if low == "true":
        return True
    if low == "false":
        return False
    if low == "none":
        return None
    return string


def _expand_includes(text, depth=0):
    if depth > 20:
        raise ValueError("#include nesting too deep")
    out = []
    for line in text.splitlines():
        m = INCLUDE_RE.match(line)
        if not m:
            out.append(line)
            continue
        spec = m.group("include")
        m_as = INCLUDE_AS_RE.match(spec)
        fname, section = (m_as.group("file"), m_as.group("as")) if m_as else (spec, None)
        with open(find_resource(fname)) as f:
            body = _expand_includes(f.read(), depth + 1)
        if section is not None:
            out.append("[%s]" % section)
        out.append(body)
    return "\n".join(out)


class PISAConfigParser(RawConfigParser):
    """RawConfigParser + ``#include``, extended interpolation, case-sensitive options."""

    def __init__(self):
        RawConfigParser.__init__(self, interpolation=ExtendedInterpolation(), empty_lines_in_values=False)

    def optionxform(self, optionstr):
        return optionstr

    def read(self, filenames, encoding=None):
        if isinstance(filenames, str):
            filenames = [filenames]
        done = []
        for fn in filenames:
            path = find_resource(fn)
            with open(path, encoding=encoding) as f:
                self.read_string(f.read(), source=path)
            done.append(path)
        return done

    def read_string(self, string, source="<string>"):
        RawConfigParser.read_file(self, StringIO(_expand_includes(string)), source)


def from_file(fname):
    cfg = PISAConfigParser()
    cfg.read(fname)
    return cfg


def interpret_param_subfields(subfields):
    """['nh','deltam31','range'] -> dict(selector='nh', pname='deltam31', attr=['range']) (:383-444)."""
    subfields = list(subfields)
    attr = None
    idx = [i for i, f in enumerate(subfields) if f in PARAM_ATTRS]
    if len(idx) > 1:
        raise ValueError("Found multiple attrs in config name %s" % subfields)
    if idx:
        attr = subfields[idx[0]:]
        subfields = subfields[:idx[0]]
    if len(subfields) == 1:
        return dict(selector=None, pname=subfields[0], attr=attr)
    if len(subfields) == 2:
        return dict(selector=subfields[0], pname=subfields[1], attr=attr)
    raise ValueError("Unable to parse param subfields %s" % subfields)


def parse_param(config, section, selector, fullname, pname, value):
    from pisa_b200.core.param import Param
    kwargs = dict(name=pname, is_fixed=True, prior=None, range=None)
    std = float("nan")
    try:
        quant, std = parse_quantity(value)
        kwargs["value"] = quant
    except ValueError:
        quant = None
        kwargs["value"] = parse_string_literal(value)
    if config.has_option(section, fullname + ".fixed"):
        kwargs["is_fixed"] = config.getboolean(section, fullname + ".fixed")
    if config.has_option(section, fullname + ".scales_as_log"):
        kwargs["scales_as_log"] = config.getboolean(section, fullname + ".scales_as_log")
    if config.has_option(section, fullname + ".tex"):
        kwargs["tex"] = config.get(section, fullname + ".tex")
    if config.has_option(section, fullname + ".range") and quant is not None:
        range_ = config.get(section, fullname + ".range")
        ns = dict(_EVAL_NS)
        ns["nominal"] = quant
        ns["sigma"] = Quantity(std, quant.units)
        range_ = range_.replace("[", "np.array([").replace("]", "], dtype=FTYPE)")
        rng = eval(range_, ns)  # pylint: disable=eval-used
        if not isinstance(rng, Quantity):
            rng = Quantity(rng, "dimensionless")
        kwargs["range"] = rng.to(quant.units)
    if config.has_option(section, fullname + ".prior"):
        kind = str(config.get(section, fullname + ".prior")).strip().lower()
        kwargs["prior"] = None if kind == "none" else dict(kind=kind)
    elif quant is not None and not np.isnan(std):
        kwargs["prior"] = dict(kind="gaussian", mean=quant, stddev=Quantity(std, quant.units))
    return Param(**kwargs)


def _parse_binning(config, binning, order):
    from pisa_b200.core.binning import MultiDimBinning, OneDimBinning
    if config["binning"].get(binning + ".split", None) is not None:
        raise NotImplementedError("VarBinning ('%s.split') is outside the scope of pisa_b200" % binning)
    dims = []
    for bin_name in order:
        kwargs = eval(config.get("binning", binning + "." + bin_name), dict(_EVAL_NS))  # pylint: disable=eval-used
        dims.append(OneDimBinning(name=bin_name, **kwargs))
    if config["binning"].get(binning + ".mask", None) is not None:
        raise NotImplementedError("bin masks are outside the scope of pisa_b200")
    return MultiDimBinning(dims, name=binning)


def parse_pipeline_config(config):
    """-> OrderedDict: 'pipeline' -> settings, (stage, service) -> constructor kwargs (:700-958)."""
    from pisa_b200.core.param import ParamSelector
    if isinstance(config, str):
        config = from_file(config)
    elif not isinstance(config, PISAConfigParser):
        raise TypeError("`config` must either be a string or PISAConfigParser. Got %s instead." % type(config))
    if not config.has_section("binning"):
        raise NoSectionError("Could not find 'binning'. Only found sections: %s" % config.sections())

    binning_dict = {}
    wanted = None
    for name in config["binning"].keys():
        if name.endswith(".order"):
            binning = name[:-len(".order")]
            order = split(config.get("binning", name))
            try:
                binning_dict[binning] = _parse_binning(config, binning, order)
            except NotImplementedError as err:
                # only an error if this binning is actually used (checked below)
                binning_dict[binning] = err

    def use_binning(key):
        b = binning_dict[key]
        if isinstance(b, Exception):
            raise b
        return b

    stage_dicts = OrderedDict()
    section = "pipeline"
    pipe = stage_dicts[section] = {}
    order = [split(x, STAGE_SEP) for x in split(config.get(section, "order"))]
    pipe["name"] = config.get(section, "name") if config.has_option(section, "name") else "none"
    if config.has_option(section, "output_binning"):
        pipe["output_binning"] = use_binning(config.get(section, "output_binning"))
        output_key = split(config.get(section, "output_key"))
        if len(output_key) == 1:
            pipe["output_key"] = output_key[0]
        elif len(output_key) == 2:
            pipe["output_key"] = tuple(output_key)
        else:
            raise ValueError("Output key should be exactly one key, or a tuple (key, error_key), but is %s" % output_key)
    else:
        pipe["output_binning"] = pipe["output_format"] = pipe["output_key"] = None
    param_selections = split(config.get(section, "param_selections")) if config.has_option(section, "param_selections") else []
    pipe["detector_name"] = config.get(section, "detector_name") if config.has_option(section, "detector_name") else None

    for stage, service in order:
        section = "%s%s%s" % (stage, STAGE_SEP, service)
        if not config.has_section(section):
            raise IOError('missing section in cfg for stage "%s" service "%s"' % (stage, service))
        service_kwargs = OrderedDict()
        selector = ParamSelector(selections=param_selections)
        service_kwargs["params"] = selector
        n_params = 0
        for fullname in config.options(section):
            value = config.get(section, fullname)
            m = PARAM_RE.match(fullname)
            if m is not None:
                n_params += 1
                info = interpret_param_subfields(m.groupdict()["subfields"].split("."))
                if info["attr"] is not None:
                    continue
                param = None
                for kw in stage_dicts.values():   # link to an identical param of an earlier stage
                    if "params" not in kw:
                        continue
                    try:
                        param = kw["params"].get(name=info["pname"], selector=info["selector"])
                    except KeyError:
                        continue
                    for a in PARAM_ATTRS:
                        if config.has_option(section, "%s.%s" % (fullname, a)):
                            raise ValueError("Parameter spec. '%s' of '%s' found in section '%s', but parameter "
                                             "exists in previous stage!" % (a, fullname, section))
                    break
                if param is None:
                    param = parse_param(config, section, info["selector"], fullname, info["pname"], value)
                selector.update(param, selector=info["selector"])
            elif value in binning_dict:
                service_kwargs[fullname] = use_binning(value)
            elif fullname in ("calc_mode", "apply_mode", "output_format"):
                service_kwargs[fullname] = parse_string_literal(value)
            elif fullname.endswith("_names"):
                service_kwargs[fullname] = split(value)
            else:
                new_value = parse_string_literal(value)
                if isinstance(new_value, str) and re.search(r"[^a-z_]units\.[a-z]+", value, flags=re.IGNORECASE):
                    try:
                        new_value = parse_quantity(value)[0]
                    except ValueError:
                        pass
                service_kwargs[fullname] = new_value
        if n_params == 0:
            service_kwargs.pop("params")
        stage_dicts[(stage, service)] = service_kwargs
    return stage_dicts
