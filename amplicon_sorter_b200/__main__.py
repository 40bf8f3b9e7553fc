from .launcher import main

main()
