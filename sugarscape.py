#! /usr/bin/python
"""Drop-in entry point: ``python sugarscape.py --conf config.json``.

Same command line (-c/--conf, -h/--help), same --conf JSON schema and same log.json output as
the reference's sugarscape.py (reference sugarscape.py:1482-1505, 1841-1972); the simulation runs
on a B200 through sugarscape_b200 (CUDA engine libssb.so).  There is no CPU fallback.
"""
import getopt
import sys

from sugarscape_b200 import config


def print_help():
    print(config.HELP_TEXT)
    sys.exit(0)


def parse_options(configuration, argv):
    try:
        args, _ = getopt.getopt(argv, "c:h:", ["conf=", "help"])
    except getopt.GetoptError as err:
        print(err)
        print_help()
    for arg, value in args:
        if arg in ("-c", "--conf"):
            if value == "":
                print("No config file provided.")
                print_help()
            config.parse_configuration(value, configuration)
        elif arg in ("-h", "--help"):
            print_help()
    return configuration


def main(argv=None):
    configuration = parse_options(config.default_configuration(), sys.argv[1:] if argv is None else argv)
    config.verify_random_seed(configuration)
    configuration = config.verify_configuration(configuration)
    if not configuration["headlessMode"]:
        sys.stderr.write("sugarscape-b200: the tkinter GUI is out of scope; running headless\n")
        configuration["headlessMode"] = True
    from sugarscape_b200.simulation import Sugarscape
    S = Sugarscape(configuration)
    if configuration["profileMode"]:
        S.engine.enable_timing(True)
    S.runSimulation(configuration["timesteps"])
    return 0


if __name__ == "__main__":
    sys.exit(main())
