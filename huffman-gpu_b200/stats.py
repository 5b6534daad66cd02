"""Series/plot sink in the file format of the reference's stats_logger (stats_logger.h:11-44,
stats_logger.cpp:13-44): one text file per (graph, series),
`<graph>__<id>_<series>.txt`, a fixed header, then one "x y" line per sample.  log_stats2 mirrors
LogStats2: time in ms against data size in MB, plus the derived `<graph>_datarate` series in GB/s
(MB * 1000 / (ms * 1024), the reference's own arithmetic)."""
import os

_series_ids = {}


def log_stats(directory, graph, series, x, y, x_quantity, y_quantity, x_unit="", y_unit="",
              x_scale="lin", y_scale="lin", series_number=0, description=""):
    key = (os.path.abspath(directory), graph, series)
    new = key not in _series_ids
    if new:
        _series_ids[key] = series_number
    path = os.path.join(directory, "%s__%d_%s.txt" % (graph, _series_ids[key], series))
    os.makedirs(directory, exist_ok=True)
    with open(path, "w" if new else "a") as f:
        if new:
            for tag, val in (("SERIES_NAME", series), ("X_AXIS_QUANTITY", x_quantity),
                             ("Y_AXIS_QUANTITY", y_quantity), ("X_AXIS_UNIT", x_unit),
                             ("Y_AXIS_UNIT", y_unit), ("X_AXIS_SCALE_TYPE", x_scale),
                             ("Y_AXIS_SCALE_TYPE", y_scale), ("DESCRIPTION", description)):
                f.write("%s\n%s\n" % (tag, val))
            f.write("__DATA__\n")
        f.write("%f %f\n" % (x, y))
    return path


def log_stats2(directory, graph, function, y_ms, x_mb, series_number=0, description=""):
    """time [ms] over data size [MB], and the data rate series the reference derives from it"""
    log_stats(directory, graph, function, x_mb, y_ms, "Data size", "Time", "MB", "ms", "log", "lin",
              series_number, description)
    log_stats(directory, graph + "_datarate", function, x_mb, (x_mb * 1000.0) / (y_ms * 1024.0), "Data size",
              "Data rate", "MB", "GB/s", "log", "lin", series_number, description)
