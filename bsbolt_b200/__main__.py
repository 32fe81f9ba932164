import sys
from bsbolt_b200.Utils.Parser import parser
from bsbolt_b200.Utils.Launcher import bsb_launch


def launch_bsb():
    arguments = parser.parse_args()
    if len(sys.argv[1:]) == 0 or arguments.subparser_name is None:
        parser.print_help()
        parser.exit()
    bsb_launch[arguments.subparser_name](arguments)


if __name__ == '__main__':
    launch_bsb()
