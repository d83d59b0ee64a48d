"""`python -m scoary_b200 -g genes.csv -t traits.csv ...` = the scoary command line."""
from .methods import main

main()
