zip = zip
range = range
map = map
