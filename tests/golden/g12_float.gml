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
    name "(-71.099, 42.3)"
    lng -71.099
    lat 42.3
  ]
  node
  [
    id 13
    name "(-71.099, 42.301)"
    lng -71.099
    lat 42.301
  ]
  node
  [
    id 14
    name "(-71.099, 42.302)"
    lng -71.099
    lat 42.302
  ]
  node
  [
    id 15
    name "(-71.099, 42.303)"
    lng -71.099
    lat 42.303
  ]
  node
  [
    id 16
    name "(-71.099, 42.304)"
    lng -71.099
    lat 42.304
  ]
  node
  [
    id 17
    name "(-71.099, 42.305)"
    lng -71.099
    lat 42.305
  ]
  node
  [
    id 18
    name "(-71.099, 42.306)"
    lng -71.099
    lat 42.306
  ]
  node
  [
    id 19
    name "(-71.099, 42.307)"
    lng -71.099
    lat 42.307
  ]
  node
  [
    id 20
    name "(-71.099, 42.308)"
    lng -71.099
    lat 42.308
  ]
  node
  [
    id 21
    name "(-71.099, 42.309)"
    lng -71.099
    lat 42.309
  ]
  node
  [
    id 22
    name "(-71.099, 42.31)"
    lng -71.099
    lat 42.31
  ]
  node
  [
    id 23
    name "(-71.099, 42.311)"
    lng -71.099
    lat 42.311
  ]
  node
  [
    id 24
    name "(-71.098, 42.3)"
    lng -71.098
    lat 42.3
  ]
  node
  [
    id 25
    name "(-71.098, 42.301)"
    lng -71.098
    lat 42.301
  ]
  node
  [
    id 26
    name "(-71.098, 42.302)"
    lng -71.098
    lat 42.302
  ]
  node
  [
    id 27
    name "(-71.098, 42.303)"
    lng -71.098
    lat 42.303
  ]
  node
  [
    id 28
    name "(-71.098, 42.304)"
    lng -71.098
    lat 42.304
  ]
  node
  [
    id 29
    name "(-71.098, 42.305)"
    lng -71.098
    lat 42.305
  ]
  node
  [
    id 30
    name "(-71.098, 42.306)"
    lng -71.098
    lat 42.306
  ]
  node
  [
    id 31
    name "(-71.098, 42.307)"
    lng -71.098
    lat 42.307
  ]
  node
  [
    id 32
    name "(-71.098, 42.308)"
    lng -71.098
    lat 42.308
  ]
  node
  [
    id 33
    name "(-71.098, 42.309)"
    lng -71.098
    lat 42.309
  ]
  node
  [
    id 34
    name "(-71.098, 42.31)"
    lng -71.098
    lat 42.31
  ]
  node
  [
    id 35
    name "(-71.098, 42.311)"
    lng -71.098
    lat 42.311
  ]
  node
  [
    id 36
    name "(-71.097, 42.3)"
    lng -71.097
    lat 42.3
  ]
  node
  [
    id 37
    name "(-71.097, 42.301)"
    lng -71.097
    lat 42.301
  ]
  node
  [
    id 38
    name "(-71.097, 42.302)"
    lng -71.097
    lat 42.302
  ]
  node
  [
    id 39
    name "(-71.097, 42.303)"
    lng -71.097
    lat 42.303
  ]
  node
  [
    id 40
    name "(-71.097, 42.304)"
    lng -71.097
    lat 42.304
  ]
  node
  [
    id 41
    name "(-71.097, 42.305)"
    lng -71.097
    lat 42.305
  ]
  node
  [
    id 42
    name "(-71.097, 42.306)"
    lng -71.097
    lat 42.306
  ]
  node
  [
    id 43
    name "(-71.097, 42.307)"
    lng -71.097
    lat 42.307
  ]
  node
  [
    id 44
    name "(-71.097, 42.308)"
    lng -71.097
    lat 42.308
  ]
  node
  [
    id 45
    name "(-71.097, 42.309)"
    lng -71.097
    lat 42.309
  ]
  node
  [
    id 46
    name "(-71.097, 42.31)"
    lng -71.097
    lat 42.31
  ]
  node
  [
    id 47
    name "(-71.097, 42.311)"
    lng -71.097
    lat 42.311
  ]
  node
  [
    id 48
    name "(-71.096, 42.3)"
    lng -71.096
    lat 42.3
  ]
  node
  [
    id 49
    name "(-71.096, 42.301)"
    lng -71.096
    lat 42.301
  ]
  node
  [
    id 50
    name "(-71.096, 42.302)"
    lng -71.096
    lat 42.302
  ]
  node
  [
    id 51
    name "(-71.096, 42.303)"
    lng -71.096
    lat 42.303
  ]
  node
  [
    id 52
    name "(-71.096, 42.304)"
    lng -71.096
    lat 42.304
  ]
  node
  [
    id 53
    name "(-71.096, 42.305)"
    lng -71.096
    lat 42.305
  ]
  node
  [
    id 54
    name "(-71.096, 42.306)"
    lng -71.096
    lat 42.306
  ]
  node
  [
    id 55
    name "(-71.096, 42.307)"
    lng -71.096
    lat 42.307
  ]
  node
  [
    id 56
    name "(-71.096, 42.308)"
    lng -71.096
    lat 42.308
  ]
  node
  [
    id 57
    name "(-71.096, 42.309)"
    lng -71.096
    lat 42.309
  ]
  node
  [
    id 58
    name "(-71.096, 42.31)"
    lng -71.096
    lat 42.31
  ]
  node
  [
    id 59
    name "(-71.096, 42.311)"
    lng -71.096
    lat 42.311
  ]
  node
  [
    id 60
    name "(-71.095, 42.3)"
    lng -71.095
    lat 42.3
  ]
  node
  [
    id 61
    name "(-71.095, 42.301)"
    lng -71.095
    lat 42.301
  ]
  node
  [
    id 62
    name "(-71.095, 42.302)"
    lng -71.095
    lat 42.302
  ]
  node
  [
    id 63
    name "(-71.095, 42.303)"
    lng -71.095
    lat 42.303
  ]
  node
  [
    id 64
    name "(-71.095, 42.304)"
    lng -71.095
    lat 42.304
  ]
  node
  [
    id 65
    name "(-71.095, 42.305)"
    lng -71.095
    lat 42.305
  ]
  node
  [
    id 66
    name "(-71.095, 42.306)"
    lng -71.095
    lat 42.306
  ]
  node
  [
    id 67
    name "(-71.095, 42.307)"
    lng -71.095
    lat 42.307
  ]
  node
  [
    id 68
    name "(-71.095, 42.308)"
    lng -71.095
    lat 42.308
  ]
  node
  [
    id 69
    name "(-71.095, 42.309)"
    lng -71.095
    lat 42.309
  ]
  node
  [
    id 70
    name "(-71.095, 42.31)"
    lng -71.095
    lat 42.31
  ]
  node
  [
    id 71
    name "(-71.095, 42.311)"
    lng -71.095
    lat 42.311
  ]
  node
  [
    id 72
    name "(-71.094, 42.3)"
    lng -71.094
    lat 42.3
  ]
  node
  [
    id 73
    name "(-71.094, 42.301)"
    lng -71.094
    lat 42.301
  ]
  node
  [
    id 74
    name "(-71.094, 42.302)"
    lng -71.094
    lat 42.302
  ]
  node
  [
    id 75
    name "(-71.094, 42.303)"
    lng -71.094
    lat 42.303
  ]
  node
  [
    id 76
    name "(-71.094, 42.304)"
    lng -71.094
    lat 42.304
  ]
  node
  [
    id 77
    name "(-71.094, 42.305)"
    lng -71.094
    lat 42.305
  ]
  node
  [
    id 78
    name "(-71.094, 42.306)"
    lng -71.094
    lat 42.306
  ]
  node
  [
    id 79
    name "(-71.094, 42.307)"
    lng -71.094
    lat 42.307
  ]
  node
  [
    id 80
    name "(-71.094, 42.308)"
    lng -71.094
    lat 42.308
  ]
  node
  [
    id 81
    name "(-71.094, 42.309)"
    lng -71.094
    lat 42.309
  ]
  node
  [
    id 82
    name "(-71.094, 42.31)"
    lng -71.094
    lat 42.31
  ]
  node
  [
    id 83
    name "(-71.094, 42.311)"
    lng -71.094
    lat 42.311
  ]
  node
  [
    id 84
    name "(-71.093, 42.3)"
    lng -71.093
    lat 42.3
  ]
  node
  [
    id 85
    name "(-71.093, 42.301)"
    lng -71.093
    lat 42.301
  ]
  node
  [
    id 86
    name "(-71.093, 42.302)"
    lng -71.093
    lat 42.302
  ]
  node
  [
    id 87
    name "(-71.093, 42.303)"
    lng -71.093
    lat 42.303
  ]
  node
  [
    id 88
    name "(-71.093, 42.304)"
    lng -71.093
    lat 42.304
  ]
  node
  [
    id 89
    name "(-71.093, 42.305)"
    lng -71.093
    lat 42.305
  ]
  node
  [
    id 90
    name "(-71.093, 42.306)"
    lng -71.093
    lat 42.306
  ]
  node
  [
    id 91
    name "(-71.093, 42.307)"
    lng -71.093
    lat 42.307
  ]
  node
  [
    id 92
    name "(-71.093, 42.308)"
    lng -71.093
    lat 42.308
  ]
  node
  [
    id 93
    name "(-71.093, 42.309)"
    lng -71.093
    lat 42.309
  ]
  node
  [
    id 94
    name "(-71.093, 42.31)"
    lng -71.093
    lat 42.31
  ]
  node
  [
    id 95
    name "(-71.093, 42.311)"
    lng -71.093
    lat 42.311
  ]
  node
  [
    id 96
    name "(-71.092, 42.3)"
    lng -71.092
    lat 42.3
  ]
  node
  [
    id 97
    name "(-71.092, 42.301)"
    lng -71.092
    lat 42.301
  ]
  node
  [
    id 98
    name "(-71.092, 42.302)"
    lng -71.092
    lat 42.302
  ]
  node
  [
    id 99
    name "(-71.092, 42.303)"
    lng -71.092
    lat 42.303
  ]
  node
  [
    id 100
    name "(-71.092, 42.304)"
    lng -71.092
    lat 42.304
  ]
  node
  [
    id 101
    name "(-71.092, 42.305)"
    lng -71.092
    lat 42.305
  ]
  node
  [
    id 102
    name "(-71.092, 42.306)"
    lng -71.092
    lat 42.306
  ]
  node
  [
    id 103
    name "(-71.092, 42.307)"
    lng -71.092
    lat 42.307
  ]
  node
  [
    id 104
    name "(-71.092, 42.308)"
    lng -71.092
    lat 42.308
  ]
  node
  [
    id 105
    name "(-71.092, 42.309)"
    lng -71.092
    lat 42.309
  ]
  node
  [
    id 106
    name "(-71.092, 42.31)"
    lng -71.092
    lat 42.31
  ]
  node
  [
    id 107
    name "(-71.092, 42.311)"
    lng -71.092
    lat 42.311
  ]
  node
  [
    id 108
    name "(-71.091, 42.3)"
    lng -71.091
    lat 42.3
  ]
  node
  [
    id 109
    name "(-71.091, 42.301)"
    lng -71.091
    lat 42.301
  ]
  node
  [
    id 110
    name "(-71.091, 42.302)"
    lng -71.091
    lat 42.302
  ]
  node
  [
    id 111
    name "(-71.091, 42.303)"
    lng -71.091
    lat 42.303
  ]
  node
  [
    id 112
    name "(-71.091, 42.304)"
    lng -71.091
    lat 42.304
  ]
  node
  [
    id 113
    name "(-71.091, 42.305)"
    lng -71.091
    lat 42.305
  ]
  node
  [
    id 114
    name "(-71.091, 42.306)"
    lng -71.091
    lat 42.306
  ]
  node
  [
    id 115
    name "(-71.091, 42.307)"
    lng -71.091
    lat 42.307
  ]
  node
  [
    id 116
    name "(-71.091, 42.308)"
    lng -71.091
    lat 42.308
  ]
  node
  [
    id 117
    name "(-71.091, 42.309)"
    lng -71.091
    lat 42.309
  ]
  node
  [
    id 118
    name "(-71.091, 42.31)"
    lng -71.091
    lat 42.31
  ]
  node
  [
    id 119
    name "(-71.091, 42.311)"
    lng -71.091
    lat 42.311
  ]
  node
  [
    id 120
    name "(-71.09, 42.3)"
    lng -71.09
    lat 42.3
  ]
  node
  [
    id 121
    name "(-71.09, 42.301)"
    lng -71.09
    lat 42.301
  ]
  node
  [
    id 122
    name "(-71.09, 42.302)"
    lng -71.09
    lat 42.302
  ]
  node
  [
    id 123
    name "(-71.09, 42.303)"
    lng -71.09
    lat 42.303
  ]
  node
  [
    id 124
    name "(-71.09, 42.304)"
    lng -71.09
    lat 42.304
  ]
  node
  [
    id 125
    name "(-71.09, 42.305)"
    lng -71.09
    lat 42.305
  ]
  node
  [
    id 126
    name "(-71.09, 42.306)"
    lng -71.09
    lat 42.306
  ]
  node
  [
    id 127
    name "(-71.09, 42.307)"
    lng -71.09
    lat 42.307
  ]
  node
  [
    id 128
    name "(-71.09, 42.308)"
    lng -71.09
    lat 42.308
  ]
  node
  [
    id 129
    name "(-71.09, 42.309)"
    lng -71.09
    lat 42.309
  ]
  node
  [
    id 130
    name "(-71.09, 42.31)"
    lng -71.09
    lat 42.31
  ]
  node
  [
    id 131
    name "(-71.09, 42.311)"
    lng -71.09
    lat 42.311
  ]
  node
  [
    id 132
    name "(-71.089, 42.3)"
    lng -71.089
    lat 42.3
  ]
  node
  [
    id 133
    name "(-71.089, 42.301)"
    lng -71.089
    lat 42.301
  ]
  node
  [
    id 134
    name "(-71.089, 42.302)"
    lng -71.089
    lat 42.302
  ]
  node
  [
    id 135
    name "(-71.089, 42.303)"
    lng -71.089
    lat 42.303
  ]
  node
  [
    id 136
    name "(-71.089, 42.304)"
    lng -71.089
    lat 42.304
  ]
  node
  [
    id 137
    name "(-71.089, 42.305)"
    lng -71.089
    lat 42.305
  ]
  node
  [
    id 138
    name "(-71.089, 42.306)"
    lng -71.089
    lat 42.306
  ]
  node
  [
    id 139
    name "(-71.089, 42.307)"
    lng -71.089
    lat 42.307
  ]
  node
  [
    id 140
    name "(-71.089, 42.308)"
    lng -71.089
    lat 42.308
  ]
  node
  [
    id 141
    name "(-71.089, 42.309)"
    lng -71.089
    lat 42.309
  ]
  node
  [
    id 142
    name "(-71.089, 42.31)"
    lng -71.089
    lat 42.31
  ]
  node
  [
    id 143
    name "(-71.089, 42.311)"
    lng -71.089
    lat 42.311
  ]
  edge
  [
    source 0
    target 1
    length 113.52815366577593
    ttmedian 12.614239296197326
    slnshift 6.307119648098663
    slnmean 1.8416790981847286
    slnsd 0.2
  ]
  edge
  [
    source 0
    target 12
    length 85.67641263798527
    ttmedian 9.519601404220586
    slnshift 4.759800702110293
    slnmean 1.5602057980622246
    slnsd 0.2
  ]
  edge
  [
    source 1
    target 2
    length 121.85946237022023
    ttmedian 13.539940263357803
    slnshift 6.769970131678901
    slnmean 1.9124966750501826
    slnsd 0.2
  ]
  edge
  [
    source 1
    target 13
    length 81.17188371763719
    ttmedian 9.019098190848576
    slnshift 4.509549095424288
    slnmean 1.506197169674748
    slnsd 0.2
  ]
  edge
  [
    source 2
    target 3
    length 138.7311204826694
    ttmedian 15.414568942518823
    slnshift 7.7072844712594115
    slnmean 2.042165916878835
    slnsd 0.2
  ]
  edge
  [
    source 3
    target 15
    length 73.36776467290198
    ttmedian 8.151973852544664
    slnshift 4.075986926272332
    slnmean 1.40511290786572
    slnsd 0.2
  ]
  edge
  [
    source 4
    target 16
    length 74.0311398791124
    ttmedian 8.225682208790268
    slnshift 4.112841104395134
    slnmean 1.414114055969849
    slnsd 0.2
  ]
  edge
  [
    source 5
    target 6
    length 119.00991822903728
    ttmedian 13.22332424767081
    slnshift 6.611662123835405
    slnmean 1.8888350782047394
    slnsd 0.2
  ]
  edge
  [
    source 5
    target 17
    length 156.0094042413918
    ttmedian 17.33437824904353
    slnshift 8.667189124521766
    slnmean 2.1595445311350545
    slnsd 0.2
  ]
  edge
  [
    source 6
    target 18
    length 81.94848889588897
    ttmedian 9.105387655098774
    slnshift 4.552693827549387
    slnmean 1.5157191077812506
    slnsd 0.2
  ]
  edge
  [
    source 7
    target 8
    length 148.06183480256837
    ttmedian 16.451314978063152
    slnshift 8.225657489031576
    slnmean 2.1072582313327533
    slnsd 0.2
  ]
  edge
  [
    source 7
    target 19
    length 146.2892847880602
    ttmedian 16.254364976451132
    slnshift 8.127182488225566
    slnmean 2.0952143060809916
    slnsd 0.2
  ]
  edge
  [
    source 8
    target 9
    length 121.8327356835048
    ttmedian 13.536970631500532
    slnshift 6.768485315750266
    slnmean 1.912277327142978
    slnsd 0.2
  ]
  edge
  [
    source 8
    target 20
    length 94.71429718867947
    ttmedian 10.523810798742163
    slnshift 5.261905399371082
    slnmean 1.6604932043818226
    slnsd 0.2
  ]
  edge
  [
    source 9
    target 10
    length 57.92535643371306
    ttmedian 6.436150714857007
    slnshift 3.2180753574285035
    slnmean 1.1687834657691278
    slnsd 0.2
  ]
  edge
  [
    source 9
    target 21
    length 67.58277101669667
    ttmedian 7.509196779632963
    slnshift 3.7545983898164814
    slnmean 1.3229813260518248
    slnsd 0.2
  ]
  edge
  [
    source 10
    target 11
    length 81.67538087261866
    ttmedian 9.07504231917985
    slnshift 4.537521159589925
    slnmean 1.512380862850603
    slnsd 0.2
  ]
  edge
  [
    source 10
    target 22
    length 113.32887192139003
    ttmedian 12.592096880154449
    slnshift 6.296048440077224
    slnmean 1.8399222048715953
    slnsd 0.2
  ]
  edge
  [
    source 11
    target 23
    length 95.27740819307158
    ttmedian 10.586378688119064
    slnshift 5.293189344059532
    slnmean 1.666420964777032
    slnsd 0.2
  ]
  edge
  [
    source 12
    target 13
    length 84.70032053350845
    ttmedian 9.411146725945382
    slnshift 4.705573362972691
    slnmean 1.548747628093626
    slnsd 0.2
  ]
  edge
  [
    source 12
    target 24
    length 102.110125984674
    ttmedian 11.345569553852666
    slnshift 5.672784776926333
    slnmean 1.7356801394838006
    slnsd 0.2
  ]
  edge
  [
    source 13
    target 14
    length 66.06002940034689
    ttmedian 7.34000326670521
    slnshift 3.670001633352605
    slnmean 1.3001921071215856
    slnsd 0.2
  ]
  edge
  [
    source 13
    target 25
    length 109.96900819760423
    ttmedian 12.218778688622692
    slnshift 6.109389344311346
    slnmean 1.8098268245409073
    slnsd 0.2
  ]
  edge
  [
    source 14
    target 15
    length 78.72824046831676
    ttmedian 8.747582274257418
    slnshift 4.373791137128709
    slnmean 1.4756301701150802
    slnsd 0.2
  ]
  edge
  [
    source 14
    target 26
    length 137.36408198490474
    ttmedian 15.262675776100528
    slnshift 7.631337888050264
    slnmean 2.0322631756721825
    slnsd 0.2
  ]
  edge
  [
    source 15
    target 16
    length 103.83468111958727
    ttmedian 11.537186791065253
    slnshift 5.768593395532626
    slnmean 1.75242827186394
    slnsd 0.2
  ]
  edge
  [
    source 15
    target 27
    length 82.58989577985575
    ttmedian 9.176655086650639
    slnshift 4.5883275433253194
    slnmean 1.5235155880278048
    slnsd 0.2
  ]
  edge
  [
    source 16
    target 17
    length 87.43669949143018
    ttmedian 9.71518883238113
    slnshift 4.857594416190565
    slnmean 1.5805433392788217
    slnsd 0.2
  ]
  edge
  [
    source 16
    target 28
    length 66.84299163848468
    ttmedian 7.426999070942742
    slnshift 3.713499535471371
    slnmean 1.3119747030353444
    slnsd 0.2
  ]
  edge
  [
    source 17
    target 29
    length 104.21096954985696
    ttmedian 11.578996616650773
    slnshift 5.789498308325387
    slnmean 1.756045639873134
    slnsd 0.2
  ]
  edge
  [
    source 18
    target 30
    length 77.67054748729531
    ttmedian 8.630060831921702
    slnshift 4.315030415960851
    slnmean 1.462104373400664
    slnsd 0.2
  ]
  edge
  [
    source 19
    target 20
    length 97.15623410368477
    ttmedian 10.79513712263164
    slnshift 5.39756856131582
    slnmean 1.6859485857468672
    slnsd 0.2
  ]
  edge
  [
    source 19
    target 31
    length 132.97748872434104
    ttmedian 14.775276524926783
    slnshift 7.387638262463391
    slnmean 1.9998080982881645
    slnsd 0.2
  ]
  edge
  [
    source 20
    target 21
    length 108.81735092176893
    ttmedian 12.090816769085437
    slnshift 6.045408384542719
    slnmean 1.7992990391947359
    slnsd 0.2
  ]
  edge
  [
    source 20
    target 32
    length 112.62332026860106
    ttmedian 12.513702252066786
    slnshift 6.256851126033393
    slnmean 1.8336770435468757
    slnsd 0.2
  ]
  edge
  [
    source 21
    target 33
    length 193.43309804614142
    ttmedian 21.49256644957127
    slnshift 10.746283224785635
    slnmean 2.3745599482579904
    slnsd 0.2
  ]
  edge
  [
    source 22
    target 34
    length 118.86008907217227
    ttmedian 13.206676563574696
    slnshift 6.603338281787348
    slnmean 1.8875753214293505
    slnsd 0.2
  ]
  edge
  [
    source 23
    target 35
    length 132.02338655132723
    ttmedian 14.669265172369693
    slnshift 7.3346325861848465
    slnmean 1.992607319840694
    slnsd 0.2
  ]
  edge
  [
    source 24
    target 25
    length 93.76101080386813
    ttmedian 10.417890089318682
    slnshift 5.208945044659341
    slnmean 1.6503773486324052
    slnsd 0.2
  ]
  edge
  [
    source 24
    target 36
    length 74.15684406597467
    ttmedian 8.239649340663853
    slnshift 4.119824670331926
    slnmean 1.4158106067112355
    slnsd 0.2
  ]
  edge
  [
    source 25
    target 26
    length 105.77423274640762
    ttmedian 11.752692527378624
    slnshift 5.876346263689312
    slnmean 1.7709351850453474
    slnsd 0.2
  ]
  edge
  [
    source 25
    target 37
    length 74.7615309254781
    ttmedian 8.306836769497567
    slnshift 4.153418384748783
    slnmean 1.4239317023504172
    slnsd 0.2
  ]
  edge
  [
    source 26
    target 27
    length 47.04908080636462
    ttmedian 5.227675645151624
    slnshift 2.613837822575812
    slnmean 0.9608195714165326
    slnsd 0.2
  ]
  edge
  [
    source 26
    target 38
    length 116.00839327734951
    ttmedian 12.889821475261057
    slnshift 6.444910737630528
    slnmean 1.8632907864318624
    slnsd 0.2
  ]
  edge
  [
    source 27
    target 28
    length 108.63769434926205
    ttmedian 12.070854927695784
    slnshift 6.035427463847892
    slnmean 1.7976466828361235
    slnsd 0.2
  ]
  edge
  [
    source 27
    target 39
    length 76.47519923547488
    ttmedian 8.49724435949721
    slnshift 4.248622179748605
    slnmean 1.446594737374086
    slnsd 0.2
  ]
  edge
  [
    source 28
    target 29
    length 138.03845848585178
    ttmedian 15.337606498427975
    slnshift 7.668803249213988
    slnmean 2.037160573115763
    slnsd 0.2
  ]
  edge
  [
    source 28
    target 40
    length 97.88465936084862
    ttmedian 10.876073262316513
    slnshift 5.438036631158257
    slnmean 1.6934180823329124
    slnsd 0.2
  ]
  edge
  [
    source 29
    target 30
    length 88.76693596588076
    ttmedian 9.862992885097862
    slnshift 4.931496442548931
    slnmean 1.5956424800386266
    slnsd 0.2
  ]
  edge
  [
    source 29
    target 41
    length 78.58977116872988
    ttmedian 8.732196796525542
    slnshift 4.366098398262771
    slnmean 1.4738697952679187
    slnsd 0.2
  ]
  edge
  [
    source 30
    target 31
    length 133.44186965393462
    ttmedian 14.826874405992735
    slnshift 7.413437202996367
    slnmean 2.003294191811669
    slnsd 0.2
  ]
  edge
  [
    source 30
    target 42
    length 89.47450038601427
    ttmedian 9.941611154001585
    slnshift 4.970805577000792
    slnmean 1.6035819149014334
    slnsd 0.2
  ]
  edge
  [
    source 31
    target 32
    length 88.03717606256477
    ttmedian 9.781908451396085
    slnshift 4.890954225698042
    slnmean 1.5873874226293614
    slnsd 0.2
  ]
  edge
  [
    source 31
    target 43
    length 88.80312801705153
    ttmedian 9.867014224116836
    slnshift 4.933507112058418
    slnmean 1.5960501168988062
    slnsd 0.2
  ]
  edge
  [
    source 32
    target 33
    length 82.04871612870116
    ttmedian 9.11652401430013
    slnshift 4.558262007150065
    slnmean 1.5169414120915674
    slnsd 0.2
  ]
  edge
  [
    source 32
    target 44
    length 65.48954660819982
    ttmedian 7.276616289799979
    slnshift 3.6383081448999897
    slnmean 1.2915187782406592
    slnsd 0.2
  ]
  edge
  [
    source 33
    target 34
    length 83.08626911191122
    ttmedian 9.231807679101246
    slnshift 4.615903839550623
    slnmean 1.5295076970023762
    slnsd 0.2
  ]
  edge
  [
    source 33
    target 45
    length 80.09851529129065
    ttmedian 8.899835032365628
    slnshift 4.449917516182814
    slnmean 1.4928855603171056
    slnsd 0.2
  ]
  edge
  [
    source 34
    target 35
    length 118.9575066165092
    ttmedian 13.21750073516769
    slnshift 6.608750367583845
    slnmean 1.8883945841883036
    slnsd 0.2
  ]
  edge
  [
    source 34
    target 46
    length 71.932229945143
    ttmedian 7.992469993904779
    slnshift 3.9962349969523894
    slnmean 1.3853526671033252
    slnsd 0.2
  ]
  edge
  [
    source 35
    target 47
    length 118.60623114922151
    ttmedian 13.178470127691279
    slnshift 6.589235063845639
    slnmean 1.8854372664881809
    slnsd 0.2
  ]
  edge
  [
    source 36
    target 37
    length 87.74409186678437
    ttmedian 9.74934354075382
    slnshift 4.87467177037691
    slnmean 1.584052773029647
    slnsd 0.2
  ]
  edge
  [
    source 36
    target 48
    length 81.49695557491444
    ttmedian 9.055217286101604
    slnshift 4.527608643050802
    slnmean 1.510193906743915
    slnsd 0.2
  ]
  edge
  [
    source 37
    target 38
    length 95.66430827237899
    ttmedian 10.629367585819887
    slnshift 5.3146837929099435
    slnmean 1.6704735166863784
    slnsd 0.2
  ]
  edge
  [
    source 37
    target 49
    length 100.81586224432182
    ttmedian 11.201762471591314
    slnshift 5.600881235795657
    slnmean 1.7229239488957
    slnsd 0.2
  ]
  edge
  [
    source 38
    target 39
    length 102.56410989248327
    ttmedian 11.39601221027592
    slnshift 5.69800610513796
    slnmean 1.740116307527926
    slnsd 0.2
  ]
  edge
  [
    source 38
    target 50
    length 115.4872925410809
    ttmedian 12.831921393453435
    slnshift 6.415960696726717
    slnmean 1.8587887447235547
    slnsd 0.2
  ]
  edge
  [
    source 39
    target 51
    length 105.40018839826627
    ttmedian 11.711132044251807
    slnshift 5.855566022125903
    slnmean 1.7673926656693315
    slnsd 0.2
  ]
  edge
  [
    source 40
    target 41
    length 67.76810072513616
    ttmedian 7.529788969459574
    slnshift 3.764894484729787
    slnmean 1.3257198355549673
    slnsd 0.2
  ]
  edge
  [
    source 40
    target 52
    length 84.08610831823185
    ttmedian 9.342900924247983
    slnshift 4.671450462123992
    slnmean 1.5414696149185032
    slnsd 0.2
  ]
  edge
  [
    source 41
    target 42
    length 150.1031097013051
    ttmedian 16.67812330014501
    slnshift 8.339061650072505
    slnmean 2.1209506980589934
    slnsd 0.2
  ]
  edge
  [
    source 41
    target 53
    length 97.02840800486794
    ttmedian 10.780934222763104
    slnshift 5.390467111381552
    slnmean 1.6846320437599
    slnsd 0.2
  ]
  edge
  [
    source 42
    target 43
    length 55.59119091408727
    ttmedian 6.176798990454142
    slnshift 3.088399495227071
    slnmean 1.1276529940109454
    slnsd 0.2
  ]
  edge
  [
    source 43
    target 44
    length 104.27008486027596
    ttmedian 11.585564984475106
    slnshift 5.792782492237553
    slnmean 1.7566127447563145
    slnsd 0.2
  ]
  edge
  [
    source 43
    target 55
    length 152.95769215584056
    ttmedian 16.99529912842673
    slnshift 8.497649564213365
    slnmean 2.139789603399801
    slnsd 0.2
  ]
  edge
  [
    source 44
    target 56
    length 133.9403951857497
    ttmedian 14.882266131749967
    slnshift 7.4411330658749835
    slnmean 2.007023131049878
    slnsd 0.2
  ]
  edge
  [
    source 45
    target 46
    length 101.3930614149896
    ttmedian 11.265895712776622
    slnshift 5.632947856388311
    slnmean 1.7286329030595622
    slnsd 0.2
  ]
  edge
  [
    source 45
    target 57
    length 105.59810022320278
    ttmedian 11.73312224702253
    slnshift 5.866561123511265
    slnmean 1.769268622903589
    slnsd 0.2
  ]
  edge
  [
    source 46
    target 47
    length 117.51475204173059
    ttmedian 13.0571946713034
    slnshift 6.5285973356517
    slnmean 1.8761921170987206
    slnsd 0.2
  ]
  edge
  [
    source 47
    target 59
    length 96.1207213930553
    ttmedian 10.680080154783923
    slnshift 5.340040077391961
    slnmean 1.675233158073521
    slnsd 0.2
  ]
  edge
  [
    source 48
    target 60
    length 88.49667088344215
    ttmedian 9.832963431493573
    slnshift 4.916481715746786
    slnmean 1.59259317627393
    slnsd 0.2
  ]
  edge
  [
    source 49
    target 50
    length 150.3853132377344
    ttmedian 16.709479248637155
    slnshift 8.354739624318578
    slnmean 2.1228289975065207
    slnsd 0.2
  ]
  edge
  [
    source 49
    target 61
    length 90.45602438330678
    ttmedian 10.050669375922975
    slnshift 5.025334687961488
    slnmean 1.6144920562966039
    slnsd 0.2
  ]
  edge
  [
    source 51
    target 52
    length 128.70335290204474
    ttmedian 14.300372544671639
    slnshift 7.150186272335819
    slnmean 1.9671384084415067
    slnsd 0.2
  ]
  edge
  [
    source 51
    target 63
    length 67.40981283567085
    ttmedian 7.489979203963427
    slnshift 3.7449896019817137
    slnmean 1.3204188404584427
    slnsd 0.2
  ]
  edge
  [
    source 52
    target 53
    length 82.1124263672839
    ttmedian 9.123602929698212
    slnshift 4.561801464849106
    slnmean 1.5177176035923416
    slnsd 0.2
  ]
  edge
  [
    source 52
    target 64
    length 74.89284166855104
    ttmedian 8.321426852061228
    slnshift 4.160713426030614
    slnmean 1.4256865562112784
    slnsd 0.2
  ]
  edge
  [
    source 53
    target 54
    length 110.80883478630759
    ttmedian 12.312092754034175
    slnshift 6.156046377017088
    slnmean 1.817434749576762
    slnsd 0.2
  ]
  edge
  [
    source 53
    target 65
    length 85.35910449742758
    ttmedian 9.48434494415862
    slnshift 4.74217247207931
    slnmean 1.5564953581499368
    slnsd 0.2
  ]
  edge
  [
    source 54
    target 55
    length 84.98336956346219
    ttmedian 9.442596618162465
    slnshift 4.721298309081233
    slnmean 1.5520838272572017
    slnsd 0.2
  ]
  edge
  [
    source 54
    target 66
    length 79.35293873770915
    ttmedian 8.816993193078794
    slnshift 4.408496596539397
    slnmean 1.48353372352228
    slnsd 0.2
  ]
  edge
  [
    source 55
    target 56
    length 103.64390903259843
    ttmedian 11.515989892510937
    slnshift 5.757994946255469
    slnmean 1.7505893144986417
    slnsd 0.2
  ]
  edge
  [
    source 56
    target 57
    length 127.61886542855612
    ttmedian 14.179873936506235
    slnshift 7.0899369682531175
    slnmean 1.9586764502722354
    slnsd 0.2
  ]
  edge
  [
    source 56
    target 68
    length 86.07606725033767
    ttmedian 9.56400747225974
    slnshift 4.78200373612987
    slnmean 1.564859650326056
    slnsd 0.2
  ]
  edge
  [
    source 57
    target 58
    length 116.38980568707777
    ttmedian 12.93220063189753
    slnshift 6.466100315948765
    slnmean 1.866573193557897
    slnsd 0.2
  ]
  edge
  [
    source 57
    target 69
    length 99.53600079938094
    ttmedian 11.05955564437566
    slnshift 5.52977782218783
    slnmean 1.7101476379075728
    slnsd 0.2
  ]
  edge
  [
    source 58
    target 59
    length 85.60813258840943
    ttmedian 9.512014732045493
    slnshift 4.756007366022747
    slnmean 1.5594085276126066
    slnsd 0.2
  ]
  edge
  [
    source 58
    target 70
    length 109.05177682210582
    ttmedian 12.116864091345091
    slnshift 6.058432045672546
    slnmean 1.8014510282648206
    slnsd 0.2
  ]
  edge
  [
    source 59
    target 71
    length 139.70639306473512
    ttmedian 15.522932562748347
    slnshift 7.761466281374173
    slnmean 2.049171270130941
    slnsd 0.2
  ]
  edge
  [
    source 60
    target 61
    length 52.13940949705198
    ttmedian 5.793267721894664
    slnshift 2.896633860947332
    slnmean 1.063549325210421
    slnsd 0.2
  ]
  edge
  [
    source 60
    target 72
    length 118.98317654391131
    ttmedian 13.220352949323479
    slnshift 6.610176474661739
    slnmean 1.8886103516406008
    slnsd 0.2
  ]
  edge
  [
    source 61
    target 62
    length 155.49880332116655
    ttmedian 17.27764481346295
    slnshift 8.638822406731475
    slnmean 2.1562662780096957
    slnsd 0.2
  ]
  edge
  [
    source 61
    target 73
    length 95.75961075726481
    ttmedian 10.639956750807201
    slnshift 5.3199783754036005
    slnmean 1.6714692385715324
    slnsd 0.2
  ]
  edge
  [
    source 62
    target 63
    length 121.0288701357324
    ttmedian 13.4476522373036
    slnshift 6.7238261186518
    slnmean 1.9056573554041163
    slnsd 0.2
  ]
  edge
  [
    source 62
    target 74
    length 127.66433692249413
    ttmedian 14.18492632472157
    slnshift 7.092463162360785
    slnmean 1.9590326938009919
    slnsd 0.2
  ]
  edge
  [
    source 63
    target 64
    length 113.81470727628886
    ttmedian 12.646078586254317
    slnshift 6.323039293127159
    slnmean 1.8441999933742887
    slnsd 0.2
  ]
  edge
  [
    source 63
    target 75
    length 106.42738534279161
    ttmedian 11.825265038087956
    slnshift 5.912632519043978
    slnmean 1.7770911669358358
    slnsd 0.2
  ]
  edge
  [
    source 64
    target 65
    length 138.5572229473057
    ttmedian 15.395246994145076
    slnshift 7.697623497072538
    slnmean 2.04091164447769
    slnsd 0.2
  ]
  edge
  [
    source 64
    target 76
    length 132.19869803677568
    ttmedian 14.688744226308408
    slnshift 7.344372113154204
    slnmean 1.9939343210370375
    slnsd 0.2
  ]
  edge
  [
    source 65
    target 77
    length 104.7638848500584
    ttmedian 11.640431650006489
    slnshift 5.8202158250032445
    slnmean 1.7613373443896674
    slnsd 0.2
  ]
  edge
  [
    source 66
    target 67
    length 118.02732763546034
    ttmedian 13.11414751505115
    slnshift 6.557073757525575
    slnmean 1.8805444298876652
    slnsd 0.2
  ]
  edge
  [
    source 66
    target 78
    length 79.75368943921274
    ttmedian 8.861521048801414
    slnshift 4.430760524400707
    slnmean 1.4885712452593711
    slnsd 0.2
  ]
  edge
  [
    source 67
    target 68
    length 96.10443956755633
    ttmedian 10.678271063061814
    slnshift 5.339135531530907
    slnmean 1.675063754386131
    slnsd 0.2
  ]
  edge
  [
    source 67
    target 79
    length 103.35157317928585
    ttmedian 11.483508131031762
    slnshift 5.741754065515881
    slnmean 1.747764749976292
    slnsd 0.2
  ]
  edge
  [
    source 68
    target 69
    length 98.90969261273034
    ttmedian 10.989965845858926
    slnshift 5.494982922929463
    slnmean 1.703835480103615
    slnsd 0.2
  ]
  edge
  [
    source 68
    target 80
    length 84.65475566865022
    ttmedian 9.406083963183358
    slnshift 4.703041981591679
    slnmean 1.5482095294788165
    slnsd 0.2
  ]
  edge
  [
    source 69
    target 70
    length 137.2075470880653
    ttmedian 15.245283009785034
    slnshift 7.622641504892517
    slnmean 2.0311229638132446
    slnsd 0.2
  ]
  edge
  [
    source 69
    target 81
    length 109.16132950815202
    ttmedian 12.129036612016892
    slnshift 6.064518306008446
    slnmean 1.8024551173142163
    slnsd 0.2
  ]
  edge
  [
    source 70
    target 71
    length 120.25182156520584
    ttmedian 13.361313507245093
    slnshift 6.6806567536225465
    slnmean 1.8992162991263737
    slnsd 0.2
  ]
  edge
  [
    source 70
    target 82
    length 72.24999978458737
    ttmedian 8.027777753843042
    slnshift 4.013888876921521
    slnmean 1.3897605661148873
    slnsd 0.2
  ]
  edge
  [
    source 71
    target 83
    length 87.99933737300847
    ttmedian 9.777704152556497
    slnshift 4.888852076278249
    slnmean 1.586957526701516
    slnsd 0.2
  ]
  edge
  [
    source 72
    target 73
    length 84.09611155547123
    ttmedian 9.344012395052358
    slnshift 4.672006197526179
    slnmean 1.5415885720503413
    slnsd 0.2
  ]
  edge
  [
    source 72
    target 84
    length 95.51573892026146
    ttmedian 10.61285988002905
    slnshift 5.306429940014525
    slnmean 1.6689192814593579
    slnsd 0.2
  ]
  edge
  [
    source 73
    target 85
    length 152.28663627891342
    ttmedian 16.920737364323713
    slnshift 8.460368682161857
    slnmean 2.1353927521203313
    slnsd 0.2
  ]
  edge
  [
    source 74
    target 86
    length 188.38751929778743
    ttmedian 20.931946588643047
    slnshift 10.465973294321524
    slnmean 2.348129356286291
    slnsd 0.2
  ]
  edge
  [
    source 75
    target 76
    length 105.15238399688972
    ttmedian 11.683598221876636
    slnshift 5.841799110938318
    slnmean 1.7650388163387518
    slnsd 0.2
  ]
  edge
  [
    source 75
    target 87
    length 109.8691293657724
    ttmedian 12.207681040641377
    slnshift 6.103840520320689
    slnmean 1.8089181665636243
    slnsd 0.2
  ]
  edge
  [
    source 76
    target 77
    length 95.2815126870351
    ttmedian 10.5868347430039
    slnshift 5.29341737150195
    slnmean 1.666464043253182
    slnsd 0.2
  ]
  edge
  [
    source 76
    target 88
    length 87.45084891522262
    ttmedian 9.71676099058029
    slnshift 4.858380495290145
    slnmean 1.5807051509566938
    slnsd 0.2
  ]
  edge
  [
    source 77
    target 78
    length 124.93098855415357
    ttmedian 13.881220950461508
    slnshift 6.940610475230754
    slnmean 1.9373897353806993
    slnsd 0.2
  ]
  edge
  [
    source 77
    target 89
    length 90.35329918390728
    ttmedian 10.039255464878586
    slnshift 5.019627732439293
    slnmean 1.6133557740692772
    slnsd 0.2
  ]
  edge
  [
    source 78
    target 79
    length 77.15454073976775
    ttmedian 8.572726748863083
    slnshift 4.2863633744315415
    slnmean 1.455438675172864
    slnsd 0.2
  ]
  edge
  [
    source 78
    target 90
    length 117.32316240911571
    ttmedian 13.03590693434619
    slnshift 6.517953467173095
    slnmean 1.874560441263328
    slnsd 0.2
  ]
  edge
  [
    source 79
    target 80
    length 96.16266791023007
    ttmedian 10.684740878914452
    slnshift 5.342370439457226
    slnmean 1.6756694569966069
    slnsd 0.2
  ]
  edge
  [
    source 79
    target 91
    length 108.6376887240173
    ttmedian 12.070854302668588
    slnshift 6.035427151334294
    slnmean 1.7976466310562609
    slnsd 0.2
  ]
  edge
  [
    source 80
    target 81
    length 108.35759623736404
    ttmedian 12.03973291526267
    slnshift 6.019866457631335
    slnmean 1.7950650759569622
    slnsd 0.2
  ]
  edge
  [
    source 80
    target 92
    length 127.79373041229587
    ttmedian 14.199303379143986
    slnshift 7.099651689571993
    slnmean 1.9600457250371346
    slnsd 0.2
  ]
  edge
  [
    source 81
    target 82
    length 117.25769299060211
    ttmedian 13.028632554511345
    slnshift 6.514316277255673
    slnmean 1.874002259138006
    slnsd 0.2
  ]
  edge
  [
    source 82
    target 83
    length 126.27799141738407
    ttmedian 14.030887935264897
    slnshift 7.015443967632448
    slnmean 1.9481139998814458
    slnsd 0.2
  ]
  edge
  [
    source 82
    target 94
    length 97.95413584736895
    ttmedian 10.883792871929883
    slnshift 5.441896435964941
    slnmean 1.6941276096700384
    slnsd 0.2
  ]
  edge
  [
    source 83
    target 95
    length 98.08123876455261
    ttmedian 10.897915418283624
    slnshift 5.448957709141812
    slnmean 1.6954243443562735
    slnsd 0.2
  ]
  edge
  [
    source 84
    target 85
    length 110.80566305118505
    ttmedian 12.311740339020561
    slnshift 6.155870169510281
    slnmean 1.8174061256811802
    slnsd 0.2
  ]
  edge
  [
    source 84
    target 96
    length 107.12426729524023
    ttmedian 11.902696366137803
    slnshift 5.951348183068902
    slnmean 1.7836177792806773
    slnsd 0.2
  ]
  edge
  [
    source 85
    target 86
    length 114.03118888908273
    ttmedian 12.67013209878697
    slnshift 6.335066049393485
    slnmean 1.8461002398266413
    slnsd 0.2
  ]
  edge
  [
    source 85
    target 97
    length 110.48140334900339
    ttmedian 12.275711483222599
    slnshift 6.1378557416112995
    slnmean 1.8144754534142364
    slnsd 0.2
  ]
  edge
  [
    source 86
    target 98
    length 173.04174450954412
    ttmedian 19.226860501060457
    slnshift 9.613430250530229
    slnmean 2.263161105237026
    slnsd 0.2
  ]
  edge
  [
    source 87
    target 88
    length 96.0961709527409
    ttmedian 10.677352328082323
    slnshift 5.3386761640411615
    slnmean 1.6749777128819117
    slnsd 0.2
  ]
  edge
  [
    source 87
    target 99
    length 113.73229391097257
    ttmedian 12.636921545663618
    slnshift 6.318460772831809
    slnmean 1.843475629893427
    slnsd 0.2
  ]
  edge
  [
    source 88
    target 89
    length 104.29722038655456
    ttmedian 11.588580042950506
    slnshift 5.794290021475253
    slnmean 1.7568729535785899
    slnsd 0.2
  ]
  edge
  [
    source 88
    target 100
    length 140.86509447572632
    ttmedian 15.651677163969591
    slnshift 7.8258385819847955
    slnmean 2.0574308977110523
    slnsd 0.2
  ]
  edge
  [
    source 89
    target 90
    length 83.34344849732841
    ttmedian 9.260383166369824
    slnshift 4.630191583184912
    slnmean 1.532598245899718
    slnsd 0.2
  ]
  edge
  [
    source 89
    target 101
    length 74.00955131959094
    ttmedian 8.22328347995455
    slnshift 4.111641739977275
    slnmean 1.41382239886531
    slnsd 0.2
  ]
  edge
  [
    source 90
    target 91
    length 135.97811553604112
    ttmedian 15.108679504004568
    slnshift 7.554339752002284
    slnmean 2.0221221997154832
    slnsd 0.2
  ]
  edge
  [
    source 91
    target 103
    length 121.51743085479521
    ttmedian 13.501936761643911
    slnshift 6.750968380821956
    slnmean 1.9096859584199628
    slnsd 0.2
  ]
  edge
  [
    source 92
    target 93
    length 105.54150371824001
    ttmedian 11.726833746471112
    slnshift 5.863416873235556
    slnmean 1.7687325178312883
    slnsd 0.2
  ]
  edge
  [
    source 92
    target 104
    length 83.28021712842377
    ttmedian 9.253357458713753
    slnshift 4.626678729356876
    slnmean 1.5318392736168496
    slnsd 0.2
  ]
  edge
  [
    source 94
    target 95
    length 138.0857486244245
    ttmedian 15.34286095826939
    slnshift 7.671430479134695
    slnmean 2.0375031011362505
    slnsd 0.2
  ]
  edge
  [
    source 94
    target 106
    length 112.05423405892057
    ttmedian 12.450470450991174
    slnshift 6.225235225495587
    slnmean 1.8286112288650587
    slnsd 0.2
  ]
  edge
  [
    source 95
    target 107
    length 86.68997196928862
    ttmedian 9.632219107698736
    slnshift 4.816109553849368
    slnmean 1.5719664556401012
    slnsd 0.2
  ]
  edge
  [
    source 96
    target 97
    length 89.01634813174869
    ttmedian 9.890705347972077
    slnshift 4.945352673986038
    slnmean 1.5984482818411878
    slnsd 0.2
  ]
  edge
  [
    source 97
    target 98
    length 85.37279433311056
    ttmedian 9.485866037012284
    slnshift 4.742933018506142
    slnmean 1.5566557246169659
    slnsd 0.2
  ]
  edge
  [
    source 98
    target 99
    length 142.39091818053936
    ttmedian 15.82121313117104
    slnshift 7.91060656558552
    slnmean 2.0682044622258275
    slnsd 0.2
  ]
  edge
  [
    source 98
    target 110
    length 95.90440169938186
    ttmedian 10.656044633264651
    slnshift 5.3280223166323255
    slnmean 1.6729801217864342
    slnsd 0.2
  ]
  edge
  [
    source 99
    target 100
    length 77.5666310678622
    ttmedian 8.6185145630958
    slnshift 4.3092572815479
    slnmean 1.460765564799245
    slnsd 0.2
  ]
  edge
  [
    source 99
    target 111
    length 86.50960698665554
    ttmedian 9.612178554072837
    slnshift 4.806089277036419
    slnmean 1.5698837133041266
    slnsd 0.2
  ]
  edge
  [
    source 100
    target 101
    length 132.3552264100692
    ttmedian 14.706136267785467
    slnshift 7.353068133892734
    slnmean 1.9951176593064375
    slnsd 0.2
  ]
  edge
  [
    source 100
    target 112
    length 116.75288392793412
    ttmedian 12.972542658659346
    slnshift 6.486271329329673
    slnmean 1.869687840085953
    slnsd 0.2
  ]
  edge
  [
    source 101
    target 102
    length 83.89557416793131
    ttmedian 9.32173046310348
    slnshift 4.66086523155174
    slnmean 1.539201102909452
    slnsd 0.2
  ]
  edge
  [
    source 101
    target 113
    length 74.80974667527391
    ttmedian 8.312194075030435
    slnshift 4.1560970375152175
    slnmean 1.4245764217409227
    slnsd 0.2
  ]
  edge
  [
    source 102
    target 103
    length 86.51674839006988
    ttmedian 9.612972043341097
    slnshift 4.8064860216705485
    slnmean 1.569966260305742
    slnsd 0.2
  ]
  edge
  [
    source 103
    target 115
    length 107.24654789808903
    ttmedian 11.91628309978767
    slnshift 5.958141549893835
    slnmean 1.78475861196353
    slnsd 0.2
  ]
  edge
  [
    source 104
    target 105
    length 90.26680092517718
    ttmedian 10.029644547241908
    slnshift 5.014822273620954
    slnmean 1.6123979818269918
    slnsd 0.2
  ]
  edge
  [
    source 105
    target 106
    length 86.7248976715256
    ttmedian 9.636099741280622
    slnshift 4.818049870640311
    slnmean 1.5723692550399626
    slnsd 0.2
  ]
  edge
  [
    source 105
    target 117
    length 100.52393270816023
    ttmedian 11.169325856462248
    slnshift 5.584662928231124
    slnmean 1.7200240776525983
    slnsd 0.2
  ]
  edge
  [
    source 106
    target 107
    length 86.5856633801567
    ttmedian 9.620629264461856
    slnshift 4.810314632230928
    slnmean 1.5707624940846134
    slnsd 0.2
  ]
  edge
  [
    source 106
    target 118
    length 75.88659801689228
    ttmedian 8.431844224099143
    slnshift 4.215922112049571
    slnmean 1.4388683366968642
    slnsd 0.2
  ]
  edge
  [
    source 107
    target 119
    length 70.29308670642087
    ttmedian 7.810342967380096
    slnshift 3.905171483690048
    slnmean 1.3623016962045562
    slnsd 0.2
  ]
  edge
  [
    source 108
    target 109
    length 79.05778714954575
    ttmedian 8.78419857217175
    slnshift 4.392099286085875
    slnmean 1.4798073100732712
    slnsd 0.2
  ]
  edge
  [
    source 108
    target 120
    length 83.48000558749702
    ttmedian 9.275556176388557
    slnshift 4.6377780881942785
    slnmean 1.5342353912449966
    slnsd 0.2
  ]
  edge
  [
    source 109
    target 110
    length 106.49002996494391
    ttmedian 11.832225551660436
    slnshift 5.916112775830218
    slnmean 1.7776796075184853
    slnsd 0.2
  ]
  edge
  [
    source 109
    target 121
    length 101.98229304124236
    ttmedian 11.331365893471373
    slnshift 5.665682946735687
    slnmean 1.734427442683414
    slnsd 0.2
  ]
  edge
  [
    source 110
    target 111
    length 81.11426997269676
    ttmedian 9.012696663632973
    slnshift 4.506348331816486
    slnmean 1.5054871430205852
    slnsd 0.2
  ]
  edge
  [
    source 110
    target 122
    length 131.48186563442
    ttmedian 14.60909618160222
    slnshift 7.30454809080111
    slnmean 1.988497180291274
    slnsd 0.2
  ]
  edge
  [
    source 111
    target 112
    length 73.81936856983748
    ttmedian 8.202152063315275
    slnshift 4.101076031657637
    slnmean 1.4112493860229045
    slnsd 0.2
  ]
  edge
  [
    source 111
    target 123
    length 127.123799176453
    ttmedian 14.124866575161445
    slnshift 7.062433287580722
    slnmean 1.9547896504216762
    slnsd 0.2
  ]
  edge
  [
    source 112
    target 113
    length 74.6805806552063
    ttmedian 8.297842295022923
    slnshift 4.148921147511461
    slnmean 1.422848335990921
    slnsd 0.2
  ]
  edge
  [
    source 112
    target 124
    length 91.98704137010066
    ttmedian 10.220782374455629
    slnshift 5.1103911872278145
    slnmean 1.6312759545591056
    slnsd 0.2
  ]
  edge
  [
    source 113
    target 114
    length 122.98055358156286
    ttmedian 13.664505953506984
    slnshift 6.832252976753492
    slnmean 1.9216544840140746
    slnsd 0.2
  ]
  edge
  [
    source 113
    target 125
    length 124.14179712019546
    ttmedian 13.793533013355052
    slnshift 6.896766506677526
    slnmean 1.9310526795448562
    slnsd 0.2
  ]
  edge
  [
    source 114
    target 115
    length 80.74684983975848
    ttmedian 8.97187220441761
    slnshift 4.485936102208805
    slnmean 1.500947192173301
    slnsd 0.2
  ]
  edge
  [
    source 114
    target 126
    length 101.81163564104567
    ttmedian 11.312403960116185
    slnshift 5.656201980058093
    slnmean 1.732752638716575
    slnsd 0.2
  ]
  edge
  [
    source 115
    target 116
    length 75.4659436535776
    ttmedian 8.38510485039751
    slnshift 4.192552425198755
    slnmean 1.4333097191567252
    slnsd 0.2
  ]
  edge
  [
    source 115
    target 127
    length 94.97301938260178
    ttmedian 10.552557709177975
    slnshift 5.276278854588988
    slnmean 1.6632210868679642
    slnsd 0.2
  ]
  edge
  [
    source 116
    target 117
    length 111.11495311583823
    ttmedian 12.346105901759804
    slnshift 6.173052950879902
    slnmean 1.8201935211944904
    slnsd 0.2
  ]
  edge
  [
    source 116
    target 128
    length 87.10355938987597
    ttmedian 9.678173265541774
    slnshift 4.839086632770887
    slnmean 1.5767259906827256
    slnsd 0.2
  ]
  edge
  [
    source 117
    target 118
    length 144.81154100730583
    ttmedian 16.09017122303398
    slnshift 8.04508561151699
    slnmean 2.0850614219684576
    slnsd 0.2
  ]
  edge
  [
    source 117
    target 129
    length 113.19949043387945
    ttmedian 12.577721159319939
    slnshift 6.2888605796599695
    slnmean 1.8387799063952903
    slnsd 0.2
  ]
  edge
  [
    source 118
    target 119
    length 118.60972934179601
    ttmedian 13.178858815755113
    slnshift 6.589429407877557
    slnmean 1.885466760225175
    slnsd 0.2
  ]
  edge
  [
    source 118
    target 130
    length 86.81904399771622
    ttmedian 9.64656044419069
    slnshift 4.823280222095345
    slnmean 1.5734542406006313
    slnsd 0.2
  ]
  edge
  [
    source 119
    target 131
    length 96.75856147334882
    ttmedian 10.750951274816536
    slnshift 5.375475637408268
    slnmean 1.6818470607792653
    slnsd 0.2
  ]
  edge
  [
    source 120
    target 121
    length 107.81196610745256
    ttmedian 11.979107345272507
    slnshift 5.989553672636253
    slnmean 1.790016897270436
    slnsd 0.2
  ]
  edge
  [
    source 120
    target 132
    length 124.22694411972124
    ttmedian 13.802993791080137
    slnshift 6.901496895540069
    slnmean 1.9317383294574533
    slnsd 0.2
  ]
  edge
  [
    source 121
    target 122
    length 119.36795765045083
    ttmedian 13.263106405605647
    slnshift 6.631553202802824
    slnmean 1.8918390456604677
    slnsd 0.2
  ]
  edge
  [
    source 121
    target 133
    length 90.36883430298191
    ttmedian 10.040981589220213
    slnshift 5.020490794610106
    slnmean 1.6135276967752918
    slnsd 0.2
  ]
  edge
  [
    source 122
    target 123
    length 81.07525764883626
    ttmedian 9.008361960981807
    slnshift 4.504180980490903
    slnmean 1.5050060721994911
    slnsd 0.2
  ]
  edge
  [
    source 122
    target 134
    length 74.37157665973623
    ttmedian 8.26350851774847
    slnshift 4.131754258874235
    slnmean 1.4187020768059166
    slnsd 0.2
  ]
  edge
  [
    source 123
    target 124
    length 170.72534162332803
    ttmedian 18.969482402592003
    slnshift 9.484741201296002
    slnmean 2.249684317964233
    slnsd 0.2
  ]
  edge
  [
    source 123
    target 135
    length 132.82857962084472
    ttmedian 14.758731068982748
    slnshift 7.379365534491374
    slnmean 1.9986876639805402
    slnsd 0.2
  ]
  edge
  [
    source 124
    target 125
    length 143.8606762857462
    ttmedian 15.984519587305133
    slnshift 7.992259793652567
    slnmean 2.0784735475312046
    slnsd 0.2
  ]
  edge
  [
    source 124
    target 136
    length 96.38619287437166
    ttmedian 10.709576986041295
    slnshift 5.3547884930206475
    slnmean 1.677991206018028
    slnsd 0.2
  ]
  edge
  [
    source 125
    target 126
    length 166.00742047314955
    ttmedian 18.445268941461062
    slnshift 9.222634470730531
    slnmean 2.221660731106769
    slnsd 0.2
  ]
  edge
  [
    source 125
    target 137
    length 113.72001590609054
    ttmedian 12.635557322898949
    slnshift 6.317778661449474
    slnmean 1.8433676687572764
    slnsd 0.2
  ]
  edge
  [
    source 126
    target 127
    length 106.1297551288153
    ttmedian 11.792195014312812
    slnshift 5.896097507156406
    slnmean 1.7742906925990667
    slnsd 0.2
  ]
  edge
  [
    source 126
    target 138
    length 87.68558826241691
    ttmedian 9.742843140268546
    slnshift 4.871421570134273
    slnmean 1.5833857980172865
    slnsd 0.2
  ]
  edge
  [
    source 127
    target 128
    length 70.2620818710581
    ttmedian 7.806897985673122
    slnshift 3.903448992836561
    slnmean 1.3618605194598026
    slnsd 0.2
  ]
  edge
  [
    source 127
    target 139
    length 118.61690836274225
    ttmedian 13.179656484749138
    slnshift 6.589828242374569
    slnmean 1.8855272848021607
    slnsd 0.2
  ]
  edge
  [
    source 128
    target 129
    length 101.88553860903598
    ttmedian 11.320615401003998
    slnshift 5.660307700501999
    slnmean 1.73347825478371
    slnsd 0.2
  ]
  edge
  [
    source 129
    target 130
    length 99.90925450524392
    ttmedian 11.101028278360436
    slnshift 5.550514139180218
    slnmean 1.7138905611578668
    slnsd 0.2
  ]
  edge
  [
    source 129
    target 141
    length 71.37455253841827
    ttmedian 7.9305058376020305
    slnshift 3.9652529188010153
    slnmean 1.3775696408963412
    slnsd 0.2
  ]
  edge
  [
    source 130
    target 131
    length 60.619877851635785
    ttmedian 6.735541983515088
    slnshift 3.367770991757544
    slnmean 1.2142510987468993
    slnsd 0.2
  ]
  edge
  [
    source 130
    target 142
    length 75.77037112495411
    ttmedian 8.418930124994901
    slnshift 4.209465062497451
    slnmean 1.4373355760812365
    slnsd 0.2
  ]
  edge
  [
    source 131
    target 143
    length 124.81818843860295
    ttmedian 13.868687604289216
    slnshift 6.934343802144608
    slnmean 1.9364864281139533
    slnsd 0.2
  ]
  edge
  [
    source 132
    target 133
    length 160.95979855938901
    ttmedian 17.884422062154336
    slnshift 8.942211031077168
    slnmean 2.190782877519577
    slnsd 0.2
  ]
  edge
  [
    source 133
    target 134
    length 84.95108798064862
    ttmedian 9.439009775627625
    slnshift 4.719504887813812
    slnmean 1.5517038974458097
    slnsd 0.2
  ]
  edge
  [
    source 134
    target 135
    length 92.09461333912205
    ttmedian 10.232734815458006
    slnshift 5.116367407729003
    slnmean 1.6324446965658659
    slnsd 0.2
  ]
  edge
  [
    source 135
    target 136
    length 67.53588472895214
    ttmedian 7.503987192105793
    slnshift 3.7519935960528965
    slnmean 1.3222873243335922
    slnsd 0.2
  ]
  edge
  [
    source 136
    target 137
    length 77.64999906473855
    ttmedian 8.627777673859839
    slnshift 4.3138888369299195
    slnmean 1.4618397796539537
    slnsd 0.2
  ]
  edge
  [
    source 137
    target 138
    length 108.32233656634516
    ttmedian 12.035815174038351
    slnshift 6.017907587019176
    slnmean 1.7947396219947622
    slnsd 0.2
  ]
  edge
  [
    source 138
    target 139
    length 82.42299294442734
    ttmedian 9.158110327158592
    slnshift 4.579055163579296
    slnmean 1.5214926806823876
    slnsd 0.2
  ]
  edge
  [
    source 139
    target 140
    length 102.58347404720773
    ttmedian 11.398163783023081
    slnshift 5.699081891511541
    slnmean 1.7403050902024289
    slnsd 0.2
  ]
  edge
  [
    source 140
    target 141
    length 136.08743754509317
    ttmedian 15.120826393899241
    slnshift 7.560413196949621
    slnmean 2.022925844379041
    slnsd 0.2
  ]
  edge
  [
    source 141
    target 142
    length 95.41809468716855
    ttmedian 10.602010520796505
    slnshift 5.301005260398252
    slnmean 1.6678964743459146
    slnsd 0.2
  ]
  edge
  [
    source 142
    target 143
    length 74.82654955881391
    ttmedian 8.314061062090435
    slnshift 4.157030531045217
    slnmean 1.4248010047335988
    slnsd 0.2
  ]
]
