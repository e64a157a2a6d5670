"""Problem libraries that define the benchmark networks (mirror of ``tensorsat.lib``)."""
