Creator "drs_abm_b200"
graph
[
  directed 0
  node
  [
    id 0
    name "(-71.1, 42.3)"
    lng -71.1
    lat 42.3
  ]
  node
  [
    id 1
    name "(-71.1, 42.301)"
    lng -71.1
    lat 42.301
  ]
  node
  [
    id 2
    name "(-71.1, 42.302)"
    lng -71.1
    lat 42.302
  ]
  node
  [
    id 3
    name "(-71.1, 42.303)"
    lng -71.1
    lat 42.303
  ]
  node
  [
    id 4
    name "(-71.1, 42.304)"
    lng -71.1
    lat 42.304
  ]
  node
  [
    id 5
    name "(-71.1, 42.305)"
    lng -71.1
    lat 42.305
  ]
  node
  [
    id 6
    name "(-71.1, 42.306)"
    lng -71.1
    lat 42.306
  ]
  node
  [
    id 7
    name "(-71.1, 42.307)"
    lng -71.1
    lat 42.307
  ]
  node
  [
    id 8
    name "(-71.1, 42.308)"
    lng -71.1
    lat 42.308
  ]
  node
  [
    id 9
    name "(-71.1, 42.309)"
    lng -71.1
    lat 42.309
  ]
  node
  [
    id 10
    name "(-71.1, 42.31)"
    lng -71.1
    lat 42.31
  ]
  node
  [
    id 11
    name "(-71.1, 42.311)"
    lng -71.1
    lat 42.311
  ]
  node
  [
    id 12
    name "(-71.1, 42.312)"
    lng -71.1
    lat 42.312
  ]
  node
  [
    id 13
    name "(-71.1, 42.313)"
    lng -71.1
    lat 42.313
  ]
  node
  [
    id 14
    name "(-71.1, 42.314)"
    lng -71.1
    lat 42.314
  ]
  node
  [
    id 15
    name "(-71.1, 42.315)"
    lng -71.1
    lat 42.315
  ]
  node
  [
    id 16
    name "(-71.1, 42.316)"
    lng -71.1
    lat 42.316
  ]
  node
  [
    id 17
    name "(-71.1, 42.317)"
    lng -71.1
    lat 42.317
  ]
  node
  [
    id 18
    name "(-71.1, 42.318)"
    lng -71.1
    lat 42.318
  ]
  node
  [
    id 19
    name "(-71.1, 42.319)"
    lng -71.1
    lat 42.319
  ]
  node
  [
    id 20
    name "(-71.099, 42.3)"
    lng -71.099
    lat 42.3
  ]
  node
  [
    id 21
    name "(-71.099, 42.301)"
    lng -71.099
    lat 42.301
  ]
  node
  [
    id 22
    name "(-71.099, 42.302)"
    lng -71.099
    lat 42.302
  ]
  node
  [
    id 23
    name "(-71.099, 42.303)"
    lng -71.099
    lat 42.303
  ]
  node
  [
    id 24
    name "(-71.099, 42.304)"
    lng -71.099
    lat 42.304
  ]
  node
  [
    id 25
    name "(-71.099, 42.305)"
    lng -71.099
    lat 42.305
  ]
  node
  [
    id 26
    name "(-71.099, 42.306)"
    lng -71.099
    lat 42.306
  ]
  node
  [
    id 27
    name "(-71.099, 42.307)"
    lng -71.099
    lat 42.307
  ]
  node
  [
    id 28
    name "(-71.099, 42.308)"
    lng -71.099
    lat 42.308
  ]
  node
  [
    id 29
    name "(-71.099, 42.309)"
    lng -71.099
    lat 42.309
  ]
  node
  [
    id 30
    name "(-71.099, 42.31)"
    lng -71.099
    lat 42.31
  ]
  node
  [
    id 31
    name "(-71.099, 42.311)"
    lng -71.099
    lat 42.311
  ]
  node
  [
    id 32
    name "(-71.099, 42.312)"
    lng -71.099
    lat 42.312
  ]
  node
  [
    id 33
    name "(-71.099, 42.313)"
    lng -71.099
    lat 42.313
  ]
  node
  [
    id 34
    name "(-71.099, 42.314)"
    lng -71.099
    lat 42.314
  ]
  node
  [
    id 35
    name "(-71.099, 42.315)"
    lng -71.099
    lat 42.315
  ]
  node
  [
    id 36
    name "(-71.099, 42.316)"
    lng -71.099
    lat 42.316
  ]
  node
  [
    id 37
    name "(-71.099, 42.317)"
    lng -71.099
    lat 42.317
  ]
  node
  [
    id 38
    name "(-71.099, 42.318)"
    lng -71.099
    lat 42.318
  ]
  node
  [
    id 39
    name "(-71.099, 42.319)"
    lng -71.099
    lat 42.319
  ]
  node
  [
    id 40
    name "(-71.098, 42.3)"
    lng -71.098
    lat 42.3
  ]
  node
  [
    id 41
    name "(-71.098, 42.301)"
    lng -71.098
    lat 42.301
  ]
  node
  [
    id 42
    name "(-71.098, 42.302)"
    lng -71.098
    lat 42.302
  ]
  node
  [
    id 43
    name "(-71.098, 42.303)"
    lng -71.098
    lat 42.303
  ]
  node
  [
    id 44
    name "(-71.098, 42.304)"
    lng -71.098
    lat 42.304
  ]
  node
  [
    id 45
    name "(-71.098, 42.305)"
    lng -71.098
    lat 42.305
  ]
  node
  [
    id 46
    name "(-71.098, 42.306)"
    lng -71.098
    lat 42.306
  ]
  node
  [
    id 47
    name "(-71.098, 42.307)"
    lng -71.098
    lat 42.307
  ]
  node
  [
    id 48
    name "(-71.098, 42.308)"
    lng -71.098
    lat 42.308
  ]
  node
  [
    id 49
    name "(-71.098, 42.309)"
    lng -71.098
    lat 42.309
  ]
  node
  [
    id 50
    name "(-71.098, 42.31)"
    lng -71.098
    lat 42.31
  ]
  node
  [
    id 51
    name "(-71.098, 42.311)"
    lng -71.098
    lat 42.311
  ]
  node
  [
    id 52
    name "(-71.098, 42.312)"
    lng -71.098
    lat 42.312
  ]
  node
  [
    id 53
    name "(-71.098, 42.313)"
    lng -71.098
    lat 42.313
  ]
  node
  [
    id 54
    name "(-71.098, 42.314)"
    lng -71.098
    lat 42.314
  ]
  node
  [
    id 55
    name "(-71.098, 42.315)"
    lng -71.098
    lat 42.315
  ]
  node
  [
    id 56
    name "(-71.098, 42.316)"
    lng -71.098
    lat 42.316
  ]
  node
  [
    id 57
    name "(-71.098, 42.317)"
    lng -71.098
    lat 42.317
  ]
  node
  [
    id 58
    name "(-71.098, 42.318)"
    lng -71.098
    lat 42.318
  ]
  node
  [
    id 59
    name "(-71.098, 42.319)"
    lng -71.098
    lat 42.319
  ]
  node
  [
    id 60
    name "(-71.097, 42.3)"
    lng -71.097
    lat 42.3
  ]
  node
  [
    id 61
    name "(-71.097, 42.301)"
    lng -71.097
    lat 42.301
  ]
  node
  [
    id 62
    name "(-71.097, 42.302)"
    lng -71.097
    lat 42.302
  ]
  node
  [
    id 63
    name "(-71.097, 42.303)"
    lng -71.097
    lat 42.303
  ]
  node
  [
    id 64
    name "(-71.097, 42.304)"
    lng -71.097
    lat 42.304
  ]
  node
  [
    id 65
    name "(-71.097, 42.305)"
    lng -71.097
    lat 42.305
  ]
  node
  [
    id 66
    name "(-71.097, 42.306)"
    lng -71.097
    lat 42.306
  ]
  node
  [
    id 67
    name "(-71.097, 42.307)"
    lng -71.097
    lat 42.307
  ]
  node
  [
    id 68
    name "(-71.097, 42.308)"
    lng -71.097
    lat 42.308
  ]
  node
  [
    id 69
    name "(-71.097, 42.309)"
    lng -71.097
    lat 42.309
  ]
  node
  [
    id 70
    name "(-71.097, 42.31)"
    lng -71.097
    lat 42.31
  ]
  node
  [
    id 71
    name "(-71.097, 42.311)"
    lng -71.097
    lat 42.311
  ]
  node
  [
    id 72
    name "(-71.097, 42.312)"
    lng -71.097
    lat 42.312
  ]
  node
  [
    id 73
    name "(-71.097, 42.313)"
    lng -71.097
    lat 42.313
  ]
  node
  [
    id 74
    name "(-71.097, 42.314)"
    lng -71.097
    lat 42.314
  ]
  node
  [
    id 75
    name "(-71.097, 42.315)"
    lng -71.097
    lat 42.315
  ]
  node
  [
    id 76
    name "(-71.097, 42.316)"
    lng -71.097
    lat 42.316
  ]
  node
  [
    id 77
    name "(-71.097, 42.317)"
    lng -71.097
    lat 42.317
  ]
  node
  [
    id 78
    name "(-71.097, 42.318)"
    lng -71.097
    lat 42.318
  ]
  node
  [
    id 79
    name "(-71.097, 42.319)"
    lng -71.097
    lat 42.319
  ]
  node
  [
    id 80
    name "(-71.096, 42.3)"
    lng -71.096
    lat 42.3
  ]
  node
  [
    id 81
    name "(-71.096, 42.301)"
    lng -71.096
    lat 42.301
  ]
  node
  [
    id 82
    name "(-71.096, 42.302)"
    lng -71.096
    lat 42.302
  ]
  node
  [
    id 83
    name "(-71.096, 42.303)"
    lng -71.096
    lat 42.303
  ]
  node
  [
    id 84
    name "(-71.096, 42.304)"
    lng -71.096
    lat 42.304
  ]
  node
  [
    id 85
    name "(-71.096, 42.305)"
    lng -71.096
    lat 42.305
  ]
  node
  [
    id 86
    name "(-71.096, 42.306)"
    lng -71.096
    lat 42.306
  ]
  node
  [
    id 87
    name "(-71.096, 42.307)"
    lng -71.096
    lat 42.307
  ]
  node
  [
    id 88
    name "(-71.096, 42.308)"
    lng -71.096
    lat 42.308
  ]
  node
  [
    id 89
    name "(-71.096, 42.309)"
    lng -71.096
    lat 42.309
  ]
  node
  [
    id 90
    name "(-71.096, 42.31)"
    lng -71.096
    lat 42.31
  ]
  node
  [
    id 91
    name "(-71.096, 42.311)"
    lng -71.096
    lat 42.311
  ]
  node
  [
    id 92
    name "(-71.096, 42.312)"
    lng -71.096
    lat 42.312
  ]
  node
  [
    id 93
    name "(-71.096, 42.313)"
    lng -71.096
    lat 42.313
  ]
  node
  [
    id 94
    name "(-71.096, 42.314)"
    lng -71.096
    lat 42.314
  ]
  node
  [
    id 95
    name "(-71.096, 42.315)"
    lng -71.096
    lat 42.315
  ]
  node
  [
    id 96
    name "(-71.096, 42.316)"
    lng -71.096
    lat 42.316
  ]
  node
  [
    id 97
    name "(-71.096, 42.317)"
    lng -71.096
    lat 42.317
  ]
  node
  [
    id 98
    name "(-71.096, 42.318)"
    lng -71.096
    lat 42.318
  ]
  node
  [
    id 99
    name "(-71.096, 42.319)"
    lng -71.096
    lat 42.319
  ]
  node
  [
    id 100
    name "(-71.095, 42.3)"
    lng -71.095
    lat 42.3
  ]
  node
  [
    id 101
    name "(-71.095, 42.301)"
    lng -71.095
    lat 42.301
  ]
  node
  [
    id 102
    name "(-71.095, 42.302)"
    lng -71.095
    lat 42.302
  ]
  node
  [
    id 103
    name "(-71.095, 42.303)"
    lng -71.095
    lat 42.303
  ]
  node
  [
    id 104
    name "(-71.095, 42.304)"
    lng -71.095
    lat 42.304
  ]
  node
  [
    id 105
    name "(-71.095, 42.305)"
    lng -71.095
    lat 42.305
  ]
  node
  [
    id 106
    name "(-71.095, 42.306)"
    lng -71.095
    lat 42.306
  ]
  node
  [
    id 107
    name "(-71.095, 42.307)"
    lng -71.095
    lat 42.307
  ]
  node
  [
    id 108
    name "(-71.095, 42.308)"
    lng -71.095
    lat 42.308
  ]
  node
  [
    id 109
    name "(-71.095, 42.309)"
    lng -71.095
    lat 42.309
  ]
  node
  [
    id 110
    name "(-71.095, 42.31)"
    lng -71.095
    lat 42.31
  ]
  node
  [
    id 111
    name "(-71.095, 42.311)"
    lng -71.095
    lat 42.311
  ]
  node
  [
    id 112
    name "(-71.095, 42.312)"
    lng -71.095
    lat 42.312
  ]
  node
  [
    id 113
    name "(-71.095, 42.313)"
    lng -71.095
    lat 42.313
  ]
  node
  [
    id 114
    name "(-71.095, 42.314)"
    lng -71.095
    lat 42.314
  ]
  node
  [
    id 115
    name "(-71.095, 42.315)"
    lng -71.095
    lat 42.315
  ]
  node
  [
    id 116
    name "(-71.095, 42.316)"
    lng -71.095
    lat 42.316
  ]
  node
  [
    id 117
    name "(-71.095, 42.317)"
    lng -71.095
    lat 42.317
  ]
  node
  [
    id 118
    name "(-71.095, 42.318)"
    lng -71.095
    lat 42.318
  ]
  node
  [
    id 119
    name "(-71.095, 42.319)"
    lng -71.095
    lat 42.319
  ]
  node
  [
    id 120
    name "(-71.094, 42.3)"
    lng -71.094
    lat 42.3
  ]
  node
  [
    id 121
    name "(-71.094, 42.301)"
    lng -71.094
    lat 42.301
  ]
  node
  [
    id 122
    name "(-71.094, 42.302)"
    lng -71.094
    lat 42.302
  ]
  node
  [
    id 123
    name "(-71.094, 42.303)"
    lng -71.094
    lat 42.303
  ]
  node
  [
    id 124
    name "(-71.094, 42.304)"
    lng -71.094
    lat 42.304
  ]
  node
  [
    id 125
    name "(-71.094, 42.305)"
    lng -71.094
    lat 42.305
  ]
  node
  [
    id 126
    name "(-71.094, 42.306)"
    lng -71.094
    lat 42.306
  ]
  node
  [
    id 127
    name "(-71.094, 42.307)"
    lng -71.094
    lat 42.307
  ]
  node
  [
    id 128
    name "(-71.094, 42.308)"
    lng -71.094
    lat 42.308
  ]
  node
  [
    id 129
    name "(-71.094, 42.309)"
    lng -71.094
    lat 42.309
  ]
  node
  [
    id 130
    name "(-71.094, 42.31)"
    lng -71.094
    lat 42.31
  ]
  node
  [
    id 131
    name "(-71.094, 42.311)"
    lng -71.094
    lat 42.311
  ]
  node
  [
    id 132
    name "(-71.094, 42.312)"
    lng -71.094
    lat 42.312
  ]
  node
  [
    id 133
    name "(-71.094, 42.313)"
    lng -71.094
    lat 42.313
  ]
  node
  [
    id 134
    name "(-71.094, 42.314)"
    lng -71.094
    lat 42.314
  ]
  node
  [
    id 135
    name "(-71.094, 42.315)"
    lng -71.094
    lat 42.315
  ]
  node
  [
    id 136
    name "(-71.094, 42.316)"
    lng -71.094
    lat 42.316
  ]
  node
  [
    id 137
    name "(-71.094, 42.317)"
    lng -71.094
    lat 42.317
  ]
  node
  [
    id 138
    name "(-71.094, 42.318)"
    lng -71.094
    lat 42.318
  ]
  node
  [
    id 139
    name "(-71.094, 42.319)"
    lng -71.094
    lat 42.319
  ]
  node
  [
    id 140
    name "(-71.093, 42.3)"
    lng -71.093
    lat 42.3
  ]
  node
  [
    id 141
    name "(-71.093, 42.301)"
    lng -71.093
    lat 42.301
  ]
  node
  [
    id 142
    name "(-71.093, 42.302)"
    lng -71.093
    lat 42.302
  ]
  node
  [
    id 143
    name "(-71.093, 42.303)"
    lng -71.093
    lat 42.303
  ]
  node
  [
    id 144
    name "(-71.093, 42.304)"
    lng -71.093
    lat 42.304
  ]
  node
  [
    id 145
    name "(-71.093, 42.305)"
    lng -71.093
    lat 42.305
  ]
  node
  [
    id 146
    name "(-71.093, 42.306)"
    lng -71.093
    lat 42.306
  ]
  node
  [
    id 147
    name "(-71.093, 42.307)"
    lng -71.093
    lat 42.307
  ]
  node
  [
    id 148
    name "(-71.093, 42.308)"
    lng -71.093
    lat 42.308
  ]
  node
  [
    id 149
    name "(-71.093, 42.309)"
    lng -71.093
    lat 42.309
  ]
  node
  [
    id 150
    name "(-71.093, 42.31)"
    lng -71.093
    lat 42.31
  ]
  node
  [
    id 151
    name "(-71.093, 42.311)"
    lng -71.093
    lat 42.311
  ]
  node
  [
    id 152
    name "(-71.093, 42.312)"
    lng -71.093
    lat 42.312
  ]
  node
  [
    id 153
    name "(-71.093, 42.313)"
    lng -71.093
    lat 42.313
  ]
  node
  [
    id 154
    name "(-71.093, 42.314)"
    lng -71.093
    lat 42.314
  ]
  node
  [
    id 155
    name "(-71.093, 42.315)"
    lng -71.093
    lat 42.315
  ]
  node
  [
    id 156
    name "(-71.093, 42.316)"
    lng -71.093
    lat 42.316
  ]
  node
  [
    id 157
    name "(-71.093, 42.317)"
    lng -71.093
    lat 42.317
  ]
  node
  [
    id 158
    name "(-71.093, 42.318)"
    lng -71.093
    lat 42.318
  ]
  node
  [
    id 159
    name "(-71.093, 42.319)"
    lng -71.093
    lat 42.319
  ]
  node
  [
    id 160
    name "(-71.092, 42.3)"
    lng -71.092
    lat 42.3
  ]
  node
  [
    id 161
    name "(-71.092, 42.301)"
    lng -71.092
    lat 42.301
  ]
  node
  [
    id 162
    name "(-71.092, 42.302)"
    lng -71.092
    lat 42.302
  ]
  node
  [
    id 163
    name "(-71.092, 42.303)"
    lng -71.092
    lat 42.303
  ]
  node
  [
    id 164
    name "(-71.092, 42.304)"
    lng -71.092
    lat 42.304
  ]
  node
  [
    id 165
    name "(-71.092, 42.305)"
    lng -71.092
    lat 42.305
  ]
  node
  [
    id 166
    name "(-71.092, 42.306)"
    lng -71.092
    lat 42.306
  ]
  node
  [
    id 167
    name "(-71.092, 42.307)"
    lng -71.092
    lat 42.307
  ]
  node
  [
    id 168
    name "(-71.092, 42.308)"
    lng -71.092
    lat 42.308
  ]
  node
  [
    id 169
    name "(-71.092, 42.309)"
    lng -71.092
    lat 42.309
  ]
  node
  [
    id 170
    name "(-71.092, 42.31)"
    lng -71.092
    lat 42.31
  ]
  node
  [
    id 171
    name "(-71.092, 42.311)"
    lng -71.092
    lat 42.311
  ]
  node
  [
    id 172
    name "(-71.092, 42.312)"
    lng -71.092
    lat 42.312
  ]
  node
  [
    id 173
    name "(-71.092, 42.313)"
    lng -71.092
    lat 42.313
  ]
  node
  [
    id 174
    name "(-71.092, 42.314)"
    lng -71.092
    lat 42.314
  ]
  node
  [
    id 175
    name "(-71.092, 42.315)"
    lng -71.092
    lat 42.315
  ]
  node
  [
    id 176
    name "(-71.092, 42.316)"
    lng -71.092
    lat 42.316
  ]
  node
  [
    id 177
    name "(-71.092, 42.317)"
    lng -71.092
    lat 42.317
  ]
  node
  [
    id 178
    name "(-71.092, 42.318)"
    lng -71.092
    lat 42.318
  ]
  node
  [
    id 179
    name "(-71.092, 42.319)"
    lng -71.092
    lat 42.319
  ]
  node
  [
    id 180
    name "(-71.091, 42.3)"
    lng -71.091
    lat 42.3
  ]
  node
  [
    id 181
    name "(-71.091, 42.301)"
    lng -71.091
    lat 42.301
  ]
  node
  [
    id 182
    name "(-71.091, 42.302)"
    lng -71.091
    lat 42.302
  ]
  node
  [
    id 183
    name "(-71.091, 42.303)"
    lng -71.091
    lat 42.303
  ]
  node
  [
    id 184
    name "(-71.091, 42.304)"
    lng -71.091
    lat 42.304
  ]
  node
  [
    id 185
    name "(-71.091, 42.305)"
    lng -71.091
    lat 42.305
  ]
  node
  [
    id 186
    name "(-71.091, 42.306)"
    lng -71.091
    lat 42.306
  ]
  node
  [
    id 187
    name "(-71.091, 42.307)"
    lng -71.091
    lat 42.307
  ]
  node
  [
    id 188
    name "(-71.091, 42.308)"
    lng -71.091
    lat 42.308
  ]
  node
  [
    id 189
    name "(-71.091, 42.309)"
    lng -71.091
    lat 42.309
  ]
  node
  [
    id 190
    name "(-71.091, 42.31)"
    lng -71.091
    lat 42.31
  ]
  node
  [
    id 191
    name "(-71.091, 42.311)"
    lng -71.091
    lat 42.311
  ]
  node
  [
    id 192
    name "(-71.091, 42.312)"
    lng -71.091
    lat 42.312
  ]
  node
  [
    id 193
    name "(-71.091, 42.313)"
    lng -71.091
    lat 42.313
  ]
  node
  [
    id 194
    name "(-71.091, 42.314)"
    lng -71.091
    lat 42.314
  ]
  node
  [
    id 195
    name "(-71.091, 42.315)"
    lng -71.091
    lat 42.315
  ]
  node
  [
    id 196
    name "(-71.091, 42.316)"
    lng -71.091
    lat 42.316
  ]
  node
  [
    id 197
    name "(-71.091, 42.317)"
    lng -71.091
    lat 42.317
  ]
  node
  [
    id 198
    name "(-71.091, 42.318)"
    lng -71.091
    lat 42.318
  ]
  node
  [
    id 199
    name "(-71.091, 42.319)"
    lng -71.091
    lat 42.319
  ]
  node
  [
    id 200
    name "(-71.09, 42.3)"
    lng -71.09
    lat 42.3
  ]
  node
  [
    id 201
    name "(-71.09, 42.301)"
    lng -71.09
    lat 42.301
  ]
  node
  [
    id 202
    name "(-71.09, 42.302)"
    lng -71.09
    lat 42.302
  ]
  node
  [
    id 203
    name "(-71.09, 42.303)"
    lng -71.09
    lat 42.303
  ]
  node
  [
    id 204
    name "(-71.09, 42.304)"
    lng -71.09
    lat 42.304
  ]
  node
  [
    id 205
    name "(-71.09, 42.305)"
    lng -71.09
    lat 42.305
  ]
  node
  [
    id 206
    name "(-71.09, 42.306)"
    lng -71.09
    lat 42.306
  ]
  node
  [
    id 207
    name "(-71.09, 42.307)"
    lng -71.09
    lat 42.307
  ]
  node
  [
    id 208
    name "(-71.09, 42.308)"
    lng -71.09
    lat 42.308
  ]
  node
  [
    id 209
    name "(-71.09, 42.309)"
    lng -71.09
    lat 42.309
  ]
  node
  [
    id 210
    name "(-71.09, 42.31)"
    lng -71.09
    lat 42.31
  ]
  node
  [
    id 211
    name "(-71.09, 42.311)"
    lng -71.09
    lat 42.311
  ]
  node
  [
    id 212
    name "(-71.09, 42.312)"
    lng -71.09
    lat 42.312
  ]
  node
  [
    id 213
    name "(-71.09, 42.313)"
    lng -71.09
    lat 42.313
  ]
  node
  [
    id 214
    name "(-71.09, 42.314)"
    lng -71.09
    lat 42.314
  ]
  node
  [
    id 215
    name "(-71.09, 42.315)"
    lng -71.09
    lat 42.315
  ]
  node
  [
    id 216
    name "(-71.09, 42.316)"
    lng -71.09
    lat 42.316
  ]
  node
  [
    id 217
    name "(-71.09, 42.317)"
    lng -71.09
    lat 42.317
  ]
  node
  [
    id 218
    name "(-71.09, 42.318)"
    lng -71.09
    lat 42.318
  ]
  node
  [
    id 219
    name "(-71.09, 42.319)"
    lng -71.09
    lat 42.319
  ]
  node
  [
    id 220
    name "(-71.089, 42.3)"
    lng -71.089
    lat 42.3
  ]
  node
  [
    id 221
    name "(-71.089, 42.301)"
    lng -71.089
    lat 42.301
  ]
  node
  [
    id 222
    name "(-71.089, 42.302)"
    lng -71.089
    lat 42.302
  ]
  node
  [
    id 223
    name "(-71.089, 42.303)"
    lng -71.089
    lat 42.303
  ]
  node
  [
    id 224
    name "(-71.089, 42.304)"
    lng -71.089
    lat 42.304
  ]
  node
  [
    id 225
    name "(-71.089, 42.305)"
    lng -71.089
    lat 42.305
  ]
  node
  [
    id 226
    name "(-71.089, 42.306)"
    lng -71.089
    lat 42.306
  ]
  node
  [
    id 227
    name "(-71.089, 42.307)"
    lng -71.089
    lat 42.307
  ]
  node
  [
    id 228
    name "(-71.089, 42.308)"
    lng -71.089
    lat 42.308
  ]
  node
  [
    id 229
    name "(-71.089, 42.309)"
    lng -71.089
    lat 42.309
  ]
  node
  [
    id 230
    name "(-71.089, 42.31)"
    lng -71.089
    lat 42.31
  ]
  node
  [
    id 231
    name "(-71.089, 42.311)"
    lng -71.089
    lat 42.311
  ]
  node
  [
    id 232
    name "(-71.089, 42.312)"
    lng -71.089
    lat 42.312
  ]
  node
  [
    id 233
    name "(-71.089, 42.313)"
    lng -71.089
    lat 42.313
  ]
  node
  [
    id 234
    name "(-71.089, 42.314)"
    lng -71.089
    lat 42.314
  ]
  node
  [
    id 235
    name "(-71.089, 42.315)"
    lng -71.089
    lat 42.315
  ]
  node
  [
    id 236
    name "(-71.089, 42.316)"
    lng -71.089
    lat 42.316
  ]
  node
  [
    id 237
    name "(-71.089, 42.317)"
    lng -71.089
    lat 42.317
  ]
  node
  [
    id 238
    name "(-71.089, 42.318)"
    lng -71.089
    lat 42.318
  ]
  node
  [
    id 239
    name "(-71.089, 42.319)"
    lng -71.089
    lat 42.319
  ]
  node
  [
    id 240
    name "(-71.088, 42.3)"
    lng -71.088
    lat 42.3
  ]
  node
  [
    id 241
    name "(-71.088, 42.301)"
    lng -71.088
    lat 42.301
  ]
  node
  [
    id 242
    name "(-71.088, 42.302)"
    lng -71.088
    lat 42.302
  ]
  node
  [
    id 243
    name "(-71.088, 42.303)"
    lng -71.088
    lat 42.303
  ]
  node
  [
    id 244
    name "(-71.088, 42.304)"
    lng -71.088
    lat 42.304
  ]
  node
  [
    id 245
    name "(-71.088, 42.305)"
    lng -71.088
    lat 42.305
  ]
  node
  [
    id 246
    name "(-71.088, 42.306)"
    lng -71.088
    lat 42.306
  ]
  node
  [
    id 247
    name "(-71.088, 42.307)"
    lng -71.088
    lat 42.307
  ]
  node
  [
    id 248
    name "(-71.088, 42.308)"
    lng -71.088
    lat 42.308
  ]
  node
  [
    id 249
    name "(-71.088, 42.309)"
    lng -71.088
    lat 42.309
  ]
  node
  [
    id 250
    name "(-71.088, 42.31)"
    lng -71.088
    lat 42.31
  ]
  node
  [
    id 251
    name "(-71.088, 42.311)"
    lng -71.088
    lat 42.311
  ]
  node
  [
    id 252
    name "(-71.088, 42.312)"
    lng -71.088
    lat 42.312
  ]
  node
  [
    id 253
    name "(-71.088, 42.313)"
    lng -71.088
    lat 42.313
  ]
  node
  [
    id 254
    name "(-71.088, 42.314)"
    lng -71.088
    lat 42.314
  ]
  node
  [
    id 255
    name "(-71.088, 42.315)"
    lng -71.088
    lat 42.315
  ]
  node
  [
    id 256
    name "(-71.088, 42.316)"
    lng -71.088
    lat 42.316
  ]
  node
  [
    id 257
    name "(-71.088, 42.317)"
    lng -71.088
    lat 42.317
  ]
  node
  [
    id 258
    name "(-71.088, 42.318)"
    lng -71.088
    lat 42.318
  ]
  node
  [
    id 259
    name "(-71.088, 42.319)"
    lng -71.088
    lat 42.319
  ]
  node
  [
    id 260
    name "(-71.087, 42.3)"
    lng -71.087
    lat 42.3
  ]
  node
  [
    id 261
    name "(-71.087, 42.301)"
    lng -71.087
    lat 42.301
  ]
  node
  [
    id 262
    name "(-71.087, 42.302)"
    lng -71.087
    lat 42.302
  ]
  node
  [
    id 263
    name "(-71.087, 42.303)"
    lng -71.087
    lat 42.303
  ]
  node
  [
    id 264
    name "(-71.087, 42.304)"
    lng -71.087
    lat 42.304
  ]
  node
  [
    id 265
    name "(-71.087, 42.305)"
    lng -71.087
    lat 42.305
  ]
  node
  [
    id 266
    name "(-71.087, 42.306)"
    lng -71.087
    lat 42.306
  ]
  node
  [
    id 267
    name "(-71.087, 42.307)"
    lng -71.087
    lat 42.307
  ]
  node
  [
    id 268
    name "(-71.087, 42.308)"
    lng -71.087
    lat 42.308
  ]
  node
  [
    id 269
    name "(-71.087, 42.309)"
    lng -71.087
    lat 42.309
  ]
  node
  [
    id 270
    name "(-71.087, 42.31)"
    lng -71.087
    lat 42.31
  ]
  node
  [
    id 271
    name "(-71.087, 42.311)"
    lng -71.087
    lat 42.311
  ]
  node
  [
    id 272
    name "(-71.087, 42.312)"
    lng -71.087
    lat 42.312
  ]
  node
  [
    id 273
    name "(-71.087, 42.313)"
    lng -71.087
    lat 42.313
  ]
  node
  [
    id 274
    name "(-71.087, 42.314)"
    lng -71.087
    lat 42.314
  ]
  node
  [
    id 275
    name "(-71.087, 42.315)"
    lng -71.087
    lat 42.315
  ]
  node
  [
    id 276
    name "(-71.087, 42.316)"
    lng -71.087
    lat 42.316
  ]
  node
  [
    id 277
    name "(-71.087, 42.317)"
    lng -71.087
    lat 42.317
  ]
  node
  [
    id 278
    name "(-71.087, 42.318)"
    lng -71.087
    lat 42.318
  ]
  node
  [
    id 279
    name "(-71.087, 42.319)"
    lng -71.087
    lat 42.319
  ]
  node
  [
    id 280
    name "(-71.086, 42.3)"
    lng -71.086
    lat 42.3
  ]
  node
  [
    id 281
    name "(-71.086, 42.301)"
    lng -71.086
    lat 42.301
  ]
  node
  [
    id 282
    name "(-71.086, 42.302)"
    lng -71.086
    lat 42.302
  ]
  node
  [
    id 283
    name "(-71.086, 42.303)"
    lng -71.086
    lat 42.303
  ]
  node
  [
    id 284
    name "(-71.086, 42.304)"
    lng -71.086
    lat 42.304
  ]
  node
  [
    id 285
    name "(-71.086, 42.305)"
    lng -71.086
    lat 42.305
  ]
  node
  [
    id 286
    name "(-71.086, 42.306)"
    lng -71.086
    lat 42.306
  ]
  node
  [
    id 287
    name "(-71.086, 42.307)"
    lng -71.086
    lat 42.307
  ]
  node
  [
    id 288
    name "(-71.086, 42.308)"
    lng -71.086
    lat 42.308
  ]
  node
  [
    id 289
    name "(-71.086, 42.309)"
    lng -71.086
    lat 42.309
  ]
  node
  [
    id 290
    name "(-71.086, 42.31)"
    lng -71.086
    lat 42.31
  ]
  node
  [
    id 291
    name "(-71.086, 42.311)"
    lng -71.086
    lat 42.311
  ]
  node
  [
    id 292
    name "(-71.086, 42.312)"
    lng -71.086
    lat 42.312
  ]
  node
  [
    id 293
    name "(-71.086, 42.313)"
    lng -71.086
    lat 42.313
  ]
  node
  [
    id 294
    name "(-71.086, 42.314)"
    lng -71.086
    lat 42.314
  ]
  node
  [
    id 295
    name "(-71.086, 42.315)"
    lng -71.086
    lat 42.315
  ]
  node
  [
    id 296
    name "(-71.086, 42.316)"
    lng -71.086
    lat 42.316
  ]
  node
  [
    id 297
    name "(-71.086, 42.317)"
    lng -71.086
    lat 42.317
  ]
  node
  [
    id 298
    name "(-71.086, 42.318)"
    lng -71.086
    lat 42.318
  ]
  node
  [
    id 299
    name "(-71.086, 42.319)"
    lng -71.086
    lat 42.319
  ]
  node
  [
    id 300
    name "(-71.085, 42.3)"
    lng -71.085
    lat 42.3
  ]
  node
  [
    id 301
    name "(-71.085, 42.301)"
    lng -71.085
    lat 42.301
  ]
  node
  [
    id 302
    name "(-71.085, 42.302)"
    lng -71.085
    lat 42.302
  ]
  node
  [
    id 303
    name "(-71.085, 42.303)"
    lng -71.085
    lat 42.303
  ]
  node
  [
    id 304
    name "(-71.085, 42.304)"
    lng -71.085
    lat 42.304
  ]
  node
  [
    id 305
    name "(-71.085, 42.305)"
    lng -71.085
    lat 42.305
  ]
  node
  [
    id 306
    name "(-71.085, 42.306)"
    lng -71.085
    lat 42.306
  ]
  node
  [
    id 307
    name "(-71.085, 42.307)"
    lng -71.085
    lat 42.307
  ]
  node
  [
    id 308
    name "(-71.085, 42.308)"
    lng -71.085
    lat 42.308
  ]
  node
  [
    id 309
    name "(-71.085, 42.309)"
    lng -71.085
    lat 42.309
  ]
  node
  [
    id 310
    name "(-71.085, 42.31)"
    lng -71.085
    lat 42.31
  ]
  node
  [
    id 311
    name "(-71.085, 42.311)"
    lng -71.085
    lat 42.311
  ]
  node
  [
    id 312
    name "(-71.085, 42.312)"
    lng -71.085
    lat 42.312
  ]
  node
  [
    id 313
    name "(-71.085, 42.313)"
    lng -71.085
    lat 42.313
  ]
  node
  [
    id 314
    name "(-71.085, 42.314)"
    lng -71.085
    lat 42.314
  ]
  node
  [
    id 315
    name "(-71.085, 42.315)"
    lng -71.085
    lat 42.315
  ]
  node
  [
    id 316
    name "(-71.085, 42.316)"
    lng -71.085
    lat 42.316
  ]
  node
  [
    id 317
    name "(-71.085, 42.317)"
    lng -71.085
    lat 42.317
  ]
  node
  [
    id 318
    name "(-71.085, 42.318)"
    lng -71.085
    lat 42.318
  ]
  node
  [
    id 319
    name "(-71.085, 42.319)"
    lng -71.085
    lat 42.319
  ]
  node
  [
    id 320
    name "(-71.084, 42.3)"
    lng -71.084
    lat 42.3
  ]
  node
  [
    id 321
    name "(-71.084, 42.301)"
    lng -71.084
    lat 42.301
  ]
  node
  [
    id 322
    name "(-71.084, 42.302)"
    lng -71.084
    lat 42.302
  ]
  node
  [
    id 323
    name "(-71.084, 42.303)"
    lng -71.084
    lat 42.303
  ]
  node
  [
    id 324
    name "(-71.084, 42.304)"
    lng -71.084
    lat 42.304
  ]
  node
  [
    id 325
    name "(-71.084, 42.305)"
    lng -71.084
    lat 42.305
  ]
  node
  [
    id 326
    name "(-71.084, 42.306)"
    lng -71.084
    lat 42.306
  ]
  node
  [
    id 327
    name "(-71.084, 42.307)"
    lng -71.084
    lat 42.307
  ]
  node
  [
    id 328
    name "(-71.084, 42.308)"
    lng -71.084
    lat 42.308
  ]
  node
  [
    id 329
    name "(-71.084, 42.309)"
    lng -71.084
    lat 42.309
  ]
  node
  [
    id 330
    name "(-71.084, 42.31)"
    lng -71.084
    lat 42.31
  ]
  node
  [
    id 331
    name "(-71.084, 42.311)"
    lng -71.084
    lat 42.311
  ]
  node
  [
    id 332
    name "(-71.084, 42.312)"
    lng -71.084
    lat 42.312
  ]
  node
  [
    id 333
    name "(-71.084, 42.313)"
    lng -71.084
    lat 42.313
  ]
  node
  [
    id 334
    name "(-71.084, 42.314)"
    lng -71.084
    lat 42.314
  ]
  node
  [
    id 335
    name "(-71.084, 42.315)"
    lng -71.084
    lat 42.315
  ]
  node
  [
    id 336
    name "(-71.084, 42.316)"
    lng -71.084
    lat 42.316
  ]
  node
  [
    id 337
    name "(-71.084, 42.317)"
    lng -71.084
    lat 42.317
  ]
  node
  [
    id 338
    name "(-71.084, 42.318)"
    lng -71.084
    lat 42.318
  ]
  node
  [
    id 339
    name "(-71.084, 42.319)"
    lng -71.084
    lat 42.319
  ]
  node
  [
    id 340
    name "(-71.083, 42.3)"
    lng -71.083
    lat 42.3
  ]
  node
  [
    id 341
    name "(-71.083, 42.301)"
    lng -71.083
    lat 42.301
  ]
  node
  [
    id 342
    name "(-71.083, 42.302)"
    lng -71.083
    lat 42.302
  ]
  node
  [
    id 343
    name "(-71.083, 42.303)"
    lng -71.083
    lat 42.303
  ]
  node
  [
    id 344
    name "(-71.083, 42.304)"
    lng -71.083
    lat 42.304
  ]
  node
  [
    id 345
    name "(-71.083, 42.305)"
    lng -71.083
    lat 42.305
  ]
  node
  [
    id 346
    name "(-71.083, 42.306)"
    lng -71.083
    lat 42.306
  ]
  node
  [
    id 347
    name "(-71.083, 42.307)"
    lng -71.083
    lat 42.307
  ]
  node
  [
    id 348
    name "(-71.083, 42.308)"
    lng -71.083
    lat 42.308
  ]
  node
  [
    id 349
    name "(-71.083, 42.309)"
    lng -71.083
    lat 42.309
  ]
  node
  [
    id 350
    name "(-71.083, 42.31)"
    lng -71.083
    lat 42.31
  ]
  node
  [
    id 351
    name "(-71.083, 42.311)"
    lng -71.083
    lat 42.311
  ]
  node
  [
    id 352
    name "(-71.083, 42.312)"
    lng -71.083
    lat 42.312
  ]
  node
  [
    id 353
    name "(-71.083, 42.313)"
    lng -71.083
    lat 42.313
  ]
  node
  [
    id 354
    name "(-71.083, 42.314)"
    lng -71.083
    lat 42.314
  ]
  node
  [
    id 355
    name "(-71.083, 42.315)"
    lng -71.083
    lat 42.315
  ]
  node
  [
    id 356
    name "(-71.083, 42.316)"
    lng -71.083
    lat 42.316
  ]
  node
  [
    id 357
    name "(-71.083, 42.317)"
    lng -71.083
    lat 42.317
  ]
  node
  [
    id 358
    name "(-71.083, 42.318)"
    lng -71.083
    lat 42.318
  ]
  node
  [
    id 359
    name "(-71.083, 42.319)"
    lng -71.083
    lat 42.319
  ]
  node
  [
    id 360
    name "(-71.082, 42.3)"
    lng -71.082
    lat 42.3
  ]
  node
  [
    id 361
    name "(-71.082, 42.301)"
    lng -71.082
    lat 42.301
  ]
  node
  [
    id 362
    name "(-71.082, 42.302)"
    lng -71.082
    lat 42.302
  ]
  node
  [
    id 363
    name "(-71.082, 42.303)"
    lng -71.082
    lat 42.303
  ]
  node
  [
    id 364
    name "(-71.082, 42.304)"
    lng -71.082
    lat 42.304
  ]
  node
  [
    id 365
    name "(-71.082, 42.305)"
    lng -71.082
    lat 42.305
  ]
  node
  [
    id 366
    name "(-71.082, 42.306)"
    lng -71.082
    lat 42.306
  ]
  node
  [
    id 367
    name "(-71.082, 42.307)"
    lng -71.082
    lat 42.307
  ]
  node
  [
    id 368
    name "(-71.082, 42.308)"
    lng -71.082
    lat 42.308
  ]
  node
  [
    id 369
    name "(-71.082, 42.309)"
    lng -71.082
    lat 42.309
  ]
  node
  [
    id 370
    name "(-71.082, 42.31)"
    lng -71.082
    lat 42.31
  ]
  node
  [
    id 371
    name "(-71.082, 42.311)"
    lng -71.082
    lat 42.311
  ]
  node
  [
    id 372
    name "(-71.082, 42.312)"
    lng -71.082
    lat 42.312
  ]
  node
  [
    id 373
    name "(-71.082, 42.313)"
    lng -71.082
    lat 42.313
  ]
  node
  [
    id 374
    name "(-71.082, 42.314)"
    lng -71.082
    lat 42.314
  ]
  node
  [
    id 375
    name "(-71.082, 42.315)"
    lng -71.082
    lat 42.315
  ]
  node
  [
    id 376
    name "(-71.082, 42.316)"
    lng -71.082
    lat 42.316
  ]
  node
  [
    id 377
    name "(-71.082, 42.317)"
    lng -71.082
    lat 42.317
  ]
  node
  [
    id 378
    name "(-71.082, 42.318)"
    lng -71.082
    lat 42.318
  ]
  node
  [
    id 379
    name "(-71.082, 42.319)"
    lng -71.082
    lat 42.319
  ]
  node
  [
    id 380
    name "(-71.081, 42.3)"
    lng -71.081
    lat 42.3
  ]
  node
  [
    id 381
    name "(-71.081, 42.301)"
    lng -71.081
    lat 42.301
  ]
  node
  [
    id 382
    name "(-71.081, 42.302)"
    lng -71.081
    lat 42.302
  ]
  node
  [
    id 383
    name "(-71.081, 42.303)"
    lng -71.081
    lat 42.303
  ]
  node
  [
    id 384
    name "(-71.081, 42.304)"
    lng -71.081
    lat 42.304
  ]
  node
  [
    id 385
    name "(-71.081, 42.305)"
    lng -71.081
    lat 42.305
  ]
  node
  [
    id 386
    name "(-71.081, 42.306)"
    lng -71.081
    lat 42.306
  ]
  node
  [
    id 387
    name "(-71.081, 42.307)"
    lng -71.081
    lat 42.307
  ]
  node
  [
    id 388
    name "(-71.081, 42.308)"
    lng -71.081
    lat 42.308
  ]
  node
  [
    id 389
    name "(-71.081, 42.309)"
    lng -71.081
    lat 42.309
  ]
  node
  [
    id 390
    name "(-71.081, 42.31)"
    lng -71.081
    lat 42.31
  ]
  node
  [
    id 391
    name "(-71.081, 42.311)"
    lng -71.081
    lat 42.311
  ]
  node
  [
    id 392
    name "(-71.081, 42.312)"
    lng -71.081
    lat 42.312
  ]
  node
  [
    id 393
    name "(-71.081, 42.313)"
    lng -71.081
    lat 42.313
  ]
  node
  [
    id 394
    name "(-71.081, 42.314)"
    lng -71.081
    lat 42.314
  ]
  node
  [
    id 395
    name "(-71.081, 42.315)"
    lng -71.081
    lat 42.315
  ]
  node
  [
    id 396
    name "(-71.081, 42.316)"
    lng -71.081
    lat 42.316
  ]
  node
  [
    id 397
    name "(-71.081, 42.317)"
    lng -71.081
    lat 42.317
  ]
  node
  [
    id 398
    name "(-71.081, 42.318)"
    lng -71.081
    lat 42.318
  ]
  node
  [
    id 399
    name "(-71.081, 42.319)"
    lng -71.081
    lat 42.319
  ]
  edge
  [
    source 0
    target 1
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 0
    target 20
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 1
    target 2
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 1
    target 21
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 2
    target 3
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 2
    target 22
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 3
    target 4
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 3
    target 23
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 4
    target 5
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 4
    target 24
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 5
    target 6
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 5
    target 25
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 6
    target 7
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 6
    target 26
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 7
    target 8
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 7
    target 27
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 8
    target 9
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 8
    target 28
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 9
    target 10
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 9
    target 29
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 10
    target 11
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 10
    target 30
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 11
    target 12
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 11
    target 31
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 12
    target 13
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 12
    target 32
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 13
    target 14
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 13
    target 33
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 14
    target 15
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 14
    target 34
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 15
    target 16
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 15
    target 35
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 16
    target 17
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 16
    target 36
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 17
    target 18
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 17
    target 37
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 18
    target 19
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 18
    target 38
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 19
    target 39
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 20
    target 21
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 20
    target 40
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 21
    target 22
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 21
    target 41
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 22
    target 23
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 22
    target 42
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 23
    target 24
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 23
    target 43
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 24
    target 25
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 24
    target 44
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 25
    target 26
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 25
    target 45
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 26
    target 27
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 26
    target 46
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 27
    target 28
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 27
    target 47
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 28
    target 29
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 28
    target 48
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 29
    target 30
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 29
    target 49
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 30
    target 31
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 30
    target 50
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 31
    target 32
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 31
    target 51
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 32
    target 33
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 32
    target 52
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 33
    target 34
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 33
    target 53
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 34
    target 35
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 34
    target 54
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 35
    target 36
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 35
    target 55
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 36
    target 37
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 36
    target 56
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 37
    target 38
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 37
    target 57
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 38
    target 39
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 38
    target 58
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 39
    target 59
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 40
    target 41
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 40
    target 60
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 41
    target 42
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 41
    target 61
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 42
    target 43
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 42
    target 62
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 43
    target 44
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 43
    target 63
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 44
    target 45
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 44
    target 64
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 45
    target 46
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 45
    target 65
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 46
    target 47
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 46
    target 66
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 47
    target 48
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 47
    target 67
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 48
    target 49
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 48
    target 68
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 49
    target 50
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 49
    target 69
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 50
    target 51
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 50
    target 70
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 51
    target 52
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 51
    target 71
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 52
    target 53
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 52
    target 72
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 53
    target 54
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 53
    target 73
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 54
    target 55
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 54
    target 74
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 55
    target 56
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 55
    target 75
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 56
    target 57
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 56
    target 76
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 57
    target 58
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 57
    target 77
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 58
    target 59
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 58
    target 78
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 59
    target 79
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 60
    target 61
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 60
    target 80
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 61
    target 62
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 61
    target 81
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 62
    target 63
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 62
    target 82
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 63
    target 64
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 63
    target 83
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 64
    target 65
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 64
    target 84
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 65
    target 66
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 65
    target 85
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 66
    target 67
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 66
    target 86
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 67
    target 68
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 67
    target 87
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 68
    target 69
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 68
    target 88
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 69
    target 70
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 69
    target 89
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 70
    target 71
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 70
    target 90
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 71
    target 72
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 71
    target 91
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 72
    target 73
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 72
    target 92
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 73
    target 74
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 73
    target 93
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 74
    target 75
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 74
    target 94
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 75
    target 76
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 75
    target 95
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 76
    target 77
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 76
    target 96
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 77
    target 78
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 77
    target 97
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 78
    target 79
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 78
    target 98
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 79
    target 99
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 80
    target 81
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 80
    target 100
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 81
    target 82
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 81
    target 101
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 82
    target 83
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 82
    target 102
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 83
    target 84
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 83
    target 103
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 84
    target 85
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 84
    target 104
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 85
    target 86
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 85
    target 105
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 86
    target 87
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 86
    target 106
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 87
    target 88
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 87
    target 107
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 88
    target 89
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 88
    target 108
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 89
    target 90
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 89
    target 109
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 90
    target 91
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 90
    target 110
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 91
    target 92
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 91
    target 111
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 92
    target 93
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 92
    target 112
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 93
    target 94
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 93
    target 113
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 94
    target 95
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 94
    target 114
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 95
    target 96
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 95
    target 115
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 96
    target 97
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 96
    target 116
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 97
    target 98
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 97
    target 117
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 98
    target 99
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 98
    target 118
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 99
    target 119
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 100
    target 101
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 100
    target 120
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 101
    target 102
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 101
    target 121
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 102
    target 103
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 102
    target 122
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 103
    target 104
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 103
    target 123
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 104
    target 105
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 104
    target 124
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 105
    target 106
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 105
    target 125
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 106
    target 107
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 106
    target 126
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 107
    target 108
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 107
    target 127
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 108
    target 109
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 108
    target 128
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 109
    target 110
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 109
    target 129
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 110
    target 111
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 110
    target 130
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 111
    target 112
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 111
    target 131
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 112
    target 113
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 112
    target 132
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 113
    target 114
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 113
    target 133
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 114
    target 115
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 114
    target 134
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 115
    target 116
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 115
    target 135
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 116
    target 117
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 116
    target 136
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 117
    target 118
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 117
    target 137
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 118
    target 119
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 118
    target 138
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 119
    target 139
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 120
    target 121
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 120
    target 140
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 121
    target 122
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 121
    target 141
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 122
    target 123
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 122
    target 142
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 123
    target 124
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 123
    target 143
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 124
    target 125
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 124
    target 144
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 125
    target 126
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 125
    target 145
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 126
    target 127
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 126
    target 146
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 127
    target 128
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 127
    target 147
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 128
    target 129
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 128
    target 148
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 129
    target 130
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 129
    target 149
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 130
    target 131
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 130
    target 150
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 131
    target 132
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 131
    target 151
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 132
    target 133
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 132
    target 152
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 133
    target 134
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 133
    target 153
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 134
    target 135
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 134
    target 154
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 135
    target 136
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 135
    target 155
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 136
    target 137
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 136
    target 156
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 137
    target 138
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 137
    target 157
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 138
    target 139
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 138
    target 158
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 139
    target 159
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 140
    target 141
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 140
    target 160
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 141
    target 142
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 141
    target 161
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 142
    target 143
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 142
    target 162
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 143
    target 144
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 143
    target 163
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 144
    target 145
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 144
    target 164
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 145
    target 146
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 145
    target 165
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 146
    target 147
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 146
    target 166
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 147
    target 148
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 147
    target 167
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 148
    target 149
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 148
    target 168
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 149
    target 150
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 149
    target 169
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 150
    target 151
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 150
    target 170
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 151
    target 152
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 151
    target 171
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 152
    target 153
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 152
    target 172
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 153
    target 154
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 153
    target 173
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 154
    target 155
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 154
    target 174
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 155
    target 156
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 155
    target 175
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 156
    target 157
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 156
    target 176
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 157
    target 158
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 157
    target 177
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 158
    target 159
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 158
    target 178
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 159
    target 179
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 160
    target 161
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 160
    target 180
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 161
    target 162
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 161
    target 181
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 162
    target 163
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 162
    target 182
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 163
    target 164
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 163
    target 183
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 164
    target 165
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 164
    target 184
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 165
    target 166
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 165
    target 185
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 166
    target 167
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 166
    target 186
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 167
    target 168
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 167
    target 187
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 168
    target 169
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 168
    target 188
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 169
    target 170
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 169
    target 189
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 170
    target 171
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 170
    target 190
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 171
    target 172
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 171
    target 191
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 172
    target 173
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 172
    target 192
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 173
    target 174
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 173
    target 193
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 174
    target 175
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 174
    target 194
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 175
    target 176
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 175
    target 195
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 176
    target 177
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 176
    target 196
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 177
    target 178
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 177
    target 197
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 178
    target 179
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 178
    target 198
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 179
    target 199
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 180
    target 181
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 180
    target 200
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 181
    target 182
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 181
    target 201
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 182
    target 183
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 182
    target 202
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 183
    target 184
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 183
    target 203
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 184
    target 185
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 184
    target 204
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 185
    target 186
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 185
    target 205
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 186
    target 187
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 186
    target 206
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 187
    target 188
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 187
    target 207
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 188
    target 189
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 188
    target 208
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 189
    target 190
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 189
    target 209
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 190
    target 191
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 190
    target 210
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 191
    target 192
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 191
    target 211
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 192
    target 193
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 192
    target 212
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 193
    target 194
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 193
    target 213
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 194
    target 195
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 194
    target 214
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 195
    target 196
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 195
    target 215
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 196
    target 197
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 196
    target 216
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 197
    target 198
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 197
    target 217
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 198
    target 199
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 198
    target 218
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 199
    target 219
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 200
    target 201
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 200
    target 220
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 201
    target 202
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 201
    target 221
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 202
    target 203
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 202
    target 222
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 203
    target 204
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 203
    target 223
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 204
    target 205
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 204
    target 224
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 205
    target 206
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 205
    target 225
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 206
    target 207
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 206
    target 226
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 207
    target 208
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 207
    target 227
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 208
    target 209
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 208
    target 228
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 209
    target 210
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 209
    target 229
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 210
    target 211
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 210
    target 230
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 211
    target 212
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 211
    target 231
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 212
    target 213
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 212
    target 232
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 213
    target 214
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 213
    target 233
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 214
    target 215
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 214
    target 234
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 215
    target 216
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 215
    target 235
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 216
    target 217
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 216
    target 236
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 217
    target 218
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 217
    target 237
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 218
    target 219
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 218
    target 238
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 219
    target 239
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 220
    target 221
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 220
    target 240
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 221
    target 222
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 221
    target 241
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 222
    target 223
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 222
    target 242
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 223
    target 224
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 223
    target 243
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 224
    target 225
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 224
    target 244
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 225
    target 226
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 225
    target 245
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 226
    target 227
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 226
    target 246
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 227
    target 228
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 227
    target 247
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 228
    target 229
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 228
    target 248
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 229
    target 230
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 229
    target 249
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 230
    target 231
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 230
    target 250
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 231
    target 232
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 231
    target 251
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 232
    target 233
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 232
    target 252
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 233
    target 234
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 233
    target 253
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 234
    target 235
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 234
    target 254
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 235
    target 236
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 235
    target 255
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 236
    target 237
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 236
    target 256
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 237
    target 238
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 237
    target 257
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 238
    target 239
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 238
    target 258
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 239
    target 259
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 240
    target 241
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 240
    target 260
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 241
    target 242
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 241
    target 261
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 242
    target 243
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 242
    target 262
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 243
    target 244
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 243
    target 263
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 244
    target 245
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 244
    target 264
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 245
    target 246
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 245
    target 265
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 246
    target 247
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 246
    target 266
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 247
    target 248
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 247
    target 267
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 248
    target 249
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 248
    target 268
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 249
    target 250
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 249
    target 269
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 250
    target 251
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 250
    target 270
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 251
    target 252
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 251
    target 271
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 252
    target 253
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 252
    target 272
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 253
    target 254
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 253
    target 273
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 254
    target 255
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 254
    target 274
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 255
    target 256
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 255
    target 275
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 256
    target 257
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 256
    target 276
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 257
    target 258
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 257
    target 277
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 258
    target 259
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 258
    target 278
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 259
    target 279
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 260
    target 261
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 260
    target 280
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 261
    target 262
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 261
    target 281
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 262
    target 263
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 262
    target 282
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 263
    target 264
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 263
    target 283
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 264
    target 265
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 264
    target 284
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 265
    target 266
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 265
    target 285
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 266
    target 267
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 266
    target 286
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 267
    target 268
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 267
    target 287
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 268
    target 269
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 268
    target 288
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 269
    target 270
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 269
    target 289
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 270
    target 271
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 270
    target 290
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 271
    target 272
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 271
    target 291
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 272
    target 273
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 272
    target 292
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 273
    target 274
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 273
    target 293
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 274
    target 275
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 274
    target 294
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 275
    target 276
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 275
    target 295
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 276
    target 277
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 276
    target 296
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 277
    target 278
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 277
    target 297
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 278
    target 279
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 278
    target 298
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 279
    target 299
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 280
    target 281
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 280
    target 300
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 281
    target 282
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 281
    target 301
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 282
    target 283
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 282
    target 302
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 283
    target 284
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 283
    target 303
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 284
    target 285
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 284
    target 304
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 285
    target 286
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 285
    target 305
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 286
    target 287
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 286
    target 306
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 287
    target 288
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 287
    target 307
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 288
    target 289
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 288
    target 308
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 289
    target 290
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 289
    target 309
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 290
    target 291
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 290
    target 310
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 291
    target 292
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 291
    target 311
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 292
    target 293
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 292
    target 312
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 293
    target 294
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 293
    target 313
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 294
    target 295
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 294
    target 314
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 295
    target 296
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 295
    target 315
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 296
    target 297
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 296
    target 316
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 297
    target 298
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 297
    target 317
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 298
    target 299
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 298
    target 318
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 299
    target 319
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 300
    target 301
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 300
    target 320
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 301
    target 302
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 301
    target 321
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 302
    target 303
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 302
    target 322
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 303
    target 304
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 303
    target 323
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 304
    target 305
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 304
    target 324
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 305
    target 306
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 305
    target 325
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 306
    target 307
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 306
    target 326
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 307
    target 308
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 307
    target 327
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 308
    target 309
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 308
    target 328
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 309
    target 310
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 309
    target 329
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 310
    target 311
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 310
    target 330
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 311
    target 312
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 311
    target 331
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 312
    target 313
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 312
    target 332
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 313
    target 314
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 313
    target 333
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 314
    target 315
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 314
    target 334
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 315
    target 316
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 315
    target 335
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 316
    target 317
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 316
    target 336
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 317
    target 318
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 317
    target 337
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 318
    target 319
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 318
    target 338
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 319
    target 339
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 320
    target 321
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 320
    target 340
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 321
    target 322
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 321
    target 341
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 322
    target 323
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 322
    target 342
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 323
    target 324
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 323
    target 343
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 324
    target 325
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 324
    target 344
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 325
    target 326
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 325
    target 345
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 326
    target 327
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 326
    target 346
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 327
    target 328
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 327
    target 347
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 328
    target 329
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 328
    target 348
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 329
    target 330
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 329
    target 349
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 330
    target 331
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 330
    target 350
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 331
    target 332
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 331
    target 351
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 332
    target 333
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 332
    target 352
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 333
    target 334
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 333
    target 353
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 334
    target 335
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 334
    target 354
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 335
    target 336
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 335
    target 355
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 336
    target 337
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 336
    target 356
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 337
    target 338
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 337
    target 357
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 338
    target 339
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 338
    target 358
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 339
    target 359
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 340
    target 341
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 340
    target 360
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 341
    target 342
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 341
    target 361
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 342
    target 343
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 342
    target 362
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 343
    target 344
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 343
    target 363
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 344
    target 345
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 344
    target 364
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 345
    target 346
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 345
    target 365
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 346
    target 347
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 346
    target 366
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 347
    target 348
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 347
    target 367
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 348
    target 349
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 348
    target 368
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 349
    target 350
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 349
    target 369
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 350
    target 351
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 350
    target 370
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 351
    target 352
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 351
    target 371
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 352
    target 353
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 352
    target 372
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 353
    target 354
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 353
    target 373
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 354
    target 355
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 354
    target 374
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 355
    target 356
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 355
    target 375
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 356
    target 357
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 356
    target 376
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 357
    target 358
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 357
    target 377
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 358
    target 359
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 358
    target 378
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 359
    target 379
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 360
    target 361
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 360
    target 380
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 361
    target 362
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 361
    target 381
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 362
    target 363
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 362
    target 382
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 363
    target 364
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 363
    target 383
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 364
    target 365
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 364
    target 384
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 365
    target 366
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 365
    target 385
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 366
    target 367
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 366
    target 386
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 367
    target 368
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 367
    target 387
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 368
    target 369
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 368
    target 388
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 369
    target 370
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 369
    target 389
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 370
    target 371
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 370
    target 390
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 371
    target 372
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 371
    target 391
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 372
    target 373
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 372
    target 392
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 373
    target 374
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 373
    target 393
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 374
    target 375
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 374
    target 394
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 375
    target 376
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 375
    target 395
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 376
    target 377
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 376
    target 396
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 377
    target 378
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 377
    target 397
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 378
    target 379
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 378
    target 398
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 379
    target 399
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 380
    target 381
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 381
    target 382
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 382
    target 383
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 383
    target 384
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 384
    target 385
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 385
    target 386
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 386
    target 387
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 387
    target 388
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 388
    target 389
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 389
    target 390
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 390
    target 391
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 391
    target 392
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 392
    target 393
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 393
    target 394
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 394
    target 395
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 395
    target 396
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 396
    target 397
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 397
    target 398
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
  edge
  [
    source 398
    target 399
    length 300.0
    ttmedian 30.0
    slnshift 15.0
    slnmean 2.70805020110221
    slnsd 0.2
  ]
]
