# tools/ holds command-line scripts (some named test_*.py) that parse sys.argv and need a GPU at import: never collect them.
collect_ignore_glob = ['*.py']
