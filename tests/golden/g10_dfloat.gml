Creator "drs_abm_b200"
graph
[
  directed 1
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
    name "(-71.099, 42.3)"
    lng -71.099
    lat 42.3
  ]
  node
  [
    id 11
    name "(-71.099, 42.301)"
    lng -71.099
    lat 42.301
  ]
  node
  [
    id 12
    name "(-71.099, 42.302)"
    lng -71.099
    lat 42.302
  ]
  node
  [
    id 13
    name "(-71.099, 42.303)"
    lng -71.099
    lat 42.303
  ]
  node
  [
    id 14
    name "(-71.099, 42.304)"
    lng -71.099
    lat 42.304
  ]
  node
  [
    id 15
    name "(-71.099, 42.305)"
    lng -71.099
    lat 42.305
  ]
  node
  [
    id 16
    name "(-71.099, 42.306)"
    lng -71.099
    lat 42.306
  ]
  node
  [
    id 17
    name "(-71.099, 42.307)"
    lng -71.099
    lat 42.307
  ]
  node
  [
    id 18
    name "(-71.099, 42.308)"
    lng -71.099
    lat 42.308
  ]
  node
  [
    id 19
    name "(-71.099, 42.309)"
    lng -71.099
    lat 42.309
  ]
  node
  [
    id 20
    name "(-71.098, 42.3)"
    lng -71.098
    lat 42.3
  ]
  node
  [
    id 21
    name "(-71.098, 42.301)"
    lng -71.098
    lat 42.301
  ]
  node
  [
    id 22
    name "(-71.098, 42.302)"
    lng -71.098
    lat 42.302
  ]
  node
  [
    id 23
    name "(-71.098, 42.303)"
    lng -71.098
    lat 42.303
  ]
  node
  [
    id 24
    name "(-71.098, 42.304)"
    lng -71.098
    lat 42.304
  ]
  node
  [
    id 25
    name "(-71.098, 42.305)"
    lng -71.098
    lat 42.305
  ]
  node
  [
    id 26
    name "(-71.098, 42.306)"
    lng -71.098
    lat 42.306
  ]
  node
  [
    id 27
    name "(-71.098, 42.307)"
    lng -71.098
    lat 42.307
  ]
  node
  [
    id 28
    name "(-71.098, 42.308)"
    lng -71.098
    lat 42.308
  ]
  node
  [
    id 29
    name "(-71.098, 42.309)"
    lng -71.098
    lat 42.309
  ]
  node
  [
    id 30
    name "(-71.097, 42.3)"
    lng -71.097
    lat 42.3
  ]
  node
  [
    id 31
    name "(-71.097, 42.301)"
    lng -71.097
    lat 42.301
  ]
  node
  [
    id 32
    name "(-71.097, 42.302)"
    lng -71.097
    lat 42.302
  ]
  node
  [
    id 33
    name "(-71.097, 42.303)"
    lng -71.097
    lat 42.303
  ]
  node
  [
    id 34
    name "(-71.097, 42.304)"
    lng -71.097
    lat 42.304
  ]
  node
  [
    id 35
    name "(-71.097, 42.305)"
    lng -71.097
    lat 42.305
  ]
  node
  [
    id 36
    name "(-71.097, 42.306)"
    lng -71.097
    lat 42.306
  ]
  node
  [
    id 37
    name "(-71.097, 42.307)"
    lng -71.097
    lat 42.307
  ]
  node
  [
    id 38
    name "(-71.097, 42.308)"
    lng -71.097
    lat 42.308
  ]
  node
  [
    id 39
    name "(-71.097, 42.309)"
    lng -71.097
    lat 42.309
  ]
  node
  [
    id 40
    name "(-71.096, 42.3)"
    lng -71.096
    lat 42.3
  ]
  node
  [
    id 41
    name "(-71.096, 42.301)"
    lng -71.096
    lat 42.301
  ]
  node
  [
    id 42
    name "(-71.096, 42.302)"
    lng -71.096
    lat 42.302
  ]
  node
  [
    id 43
    name "(-71.096, 42.303)"
    lng -71.096
    lat 42.303
  ]
  node
  [
    id 44
    name "(-71.096, 42.304)"
    lng -71.096
    lat 42.304
  ]
  node
  [
    id 45
    name "(-71.096, 42.305)"
    lng -71.096
    lat 42.305
  ]
  node
  [
    id 46
    name "(-71.096, 42.306)"
    lng -71.096
    lat 42.306
  ]
  node
  [
    id 47
    name "(-71.096, 42.307)"
    lng -71.096
    lat 42.307
  ]
  node
  [
    id 48
    name "(-71.096, 42.308)"
    lng -71.096
    lat 42.308
  ]
  node
  [
    id 49
    name "(-71.096, 42.309)"
    lng -71.096
    lat 42.309
  ]
  node
  [
    id 50
    name "(-71.095, 42.3)"
    lng -71.095
    lat 42.3
  ]
  node
  [
    id 51
    name "(-71.095, 42.301)"
    lng -71.095
    lat 42.301
  ]
  node
  [
    id 52
    name "(-71.095, 42.302)"
    lng -71.095
    lat 42.302
  ]
  node
  [
    id 53
    name "(-71.095, 42.303)"
    lng -71.095
    lat 42.303
  ]
  node
  [
    id 54
    name "(-71.095, 42.304)"
    lng -71.095
    lat 42.304
  ]
  node
  [
    id 55
    name "(-71.095, 42.305)"
    lng -71.095
    lat 42.305
  ]
  node
  [
    id 56
    name "(-71.095, 42.306)"
    lng -71.095
    lat 42.306
  ]
  node
  [
    id 57
    name "(-71.095, 42.307)"
    lng -71.095
    lat 42.307
  ]
  node
  [
    id 58
    name "(-71.095, 42.308)"
    lng -71.095
    lat 42.308
  ]
  node
  [
    id 59
    name "(-71.095, 42.309)"
    lng -71.095
    lat 42.309
  ]
  node
  [
    id 60
    name "(-71.094, 42.3)"
    lng -71.094
    lat 42.3
  ]
  node
  [
    id 61
    name "(-71.094, 42.301)"
    lng -71.094
    lat 42.301
  ]
  node
  [
    id 62
    name "(-71.094, 42.302)"
    lng -71.094
    lat 42.302
  ]
  node
  [
    id 63
    name "(-71.094, 42.303)"
    lng -71.094
    lat 42.303
  ]
  node
  [
    id 64
    name "(-71.094, 42.304)"
    lng -71.094
    lat 42.304
  ]
  node
  [
    id 65
    name "(-71.094, 42.305)"
    lng -71.094
    lat 42.305
  ]
  node
  [
    id 66
    name "(-71.094, 42.306)"
    lng -71.094
    lat 42.306
  ]
  node
  [
    id 67
    name "(-71.094, 42.307)"
    lng -71.094
    lat 42.307
  ]
  node
  [
    id 68
    name "(-71.094, 42.308)"
    lng -71.094
    lat 42.308
  ]
  node
  [
    id 69
    name "(-71.094, 42.309)"
    lng -71.094
    lat 42.309
  ]
  node
  [
    id 70
    name "(-71.093, 42.3)"
    lng -71.093
    lat 42.3
  ]
  node
  [
    id 71
    name "(-71.093, 42.301)"
    lng -71.093
    lat 42.301
  ]
  node
  [
    id 72
    name "(-71.093, 42.302)"
    lng -71.093
    lat 42.302
  ]
  node
  [
    id 73
    name "(-71.093, 42.303)"
    lng -71.093
    lat 42.303
  ]
  node
  [
    id 74
    name "(-71.093, 42.304)"
    lng -71.093
    lat 42.304
  ]
  node
  [
    id 75
    name "(-71.093, 42.305)"
    lng -71.093
    lat 42.305
  ]
  node
  [
    id 76
    name "(-71.093, 42.306)"
    lng -71.093
    lat 42.306
  ]
  node
  [
    id 77
    name "(-71.093, 42.307)"
    lng -71.093
    lat 42.307
  ]
  node
  [
    id 78
    name "(-71.093, 42.308)"
    lng -71.093
    lat 42.308
  ]
  node
  [
    id 79
    name "(-71.093, 42.309)"
    lng -71.093
    lat 42.309
  ]
  node
  [
    id 80
    name "(-71.092, 42.3)"
    lng -71.092
    lat 42.3
  ]
  node
  [
    id 81
    name "(-71.092, 42.301)"
    lng -71.092
    lat 42.301
  ]
  node
  [
    id 82
    name "(-71.092, 42.302)"
    lng -71.092
    lat 42.302
  ]
  node
  [
    id 83
    name "(-71.092, 42.303)"
    lng -71.092
    lat 42.303
  ]
  node
  [
    id 84
    name "(-71.092, 42.304)"
    lng -71.092
    lat 42.304
  ]
  node
  [
    id 85
    name "(-71.092, 42.305)"
    lng -71.092
    lat 42.305
  ]
  node
  [
    id 86
    name "(-71.092, 42.306)"
    lng -71.092
    lat 42.306
  ]
  node
  [
    id 87
    name "(-71.092, 42.307)"
    lng -71.092
    lat 42.307
  ]
  node
  [
    id 88
    name "(-71.092, 42.308)"
    lng -71.092
    lat 42.308
  ]
  node
  [
    id 89
    name "(-71.092, 42.309)"
    lng -71.092
    lat 42.309
  ]
  node
  [
    id 90
    name "(-71.091, 42.3)"
    lng -71.091
    lat 42.3
  ]
  node
  [
    id 91
    name "(-71.091, 42.301)"
    lng -71.091
    lat 42.301
  ]
  node
  [
    id 92
    name "(-71.091, 42.302)"
    lng -71.091
    lat 42.302
  ]
  node
  [
    id 93
    name "(-71.091, 42.303)"
    lng -71.091
    lat 42.303
  ]
  node
  [
    id 94
    name "(-71.091, 42.304)"
    lng -71.091
    lat 42.304
  ]
  node
  [
    id 95
    name "(-71.091, 42.305)"
    lng -71.091
    lat 42.305
  ]
  node
  [
    id 96
    name "(-71.091, 42.306)"
    lng -71.091
    lat 42.306
  ]
  node
  [
    id 97
    name "(-71.091, 42.307)"
    lng -71.091
    lat 42.307
  ]
  node
  [
    id 98
    name "(-71.091, 42.308)"
    lng -71.091
    lat 42.308
  ]
  node
  [
    id 99
    name "(-71.091, 42.309)"
    lng -71.091
    lat 42.309
  ]
  edge
  [
    source 0
    target 1
    length 139.50030840122844
    ttmedian 15.50003426680316
    slnshift 7.75001713340158
    slnmean 2.047695054124306
    slnsd 0.2
  ]
  edge
  [
    source 0
    target 10
    length 119.58051732033445
    ttmedian 13.286724146703827
    slnshift 6.643362073351914
    slnmean 1.893618171691528
    slnsd 0.2
  ]
  edge
  [
    source 1
    target 2
    length 67.95325980026365
    ttmedian 7.550362200029294
    slnshift 3.775181100014647
    slnmean 1.3284483550641095
    slnsd 0.2
  ]
  edge
  [
    source 1
    target 11
    length 99.79062325134221
    ttmedian 11.087847027926912
    slnshift 5.543923513963456
    slnmean 1.7127024656097962
    slnsd 0.2
  ]
  edge
  [
    source 1
    target 0
    length 116.80480163767184
    ttmedian 12.978311293074649
    slnshift 6.489155646537324
    slnmean 1.8701324215645467
    slnsd 0.2
  ]
  edge
  [
    source 2
    target 3
    length 83.5252345022415
    ttmedian 9.280581611360168
    slnshift 4.640290805680084
    slnmean 1.5347770379122025
    slnsd 0.2
  ]
  edge
  [
    source 2
    target 12
    length 106.86304725054396
    ttmedian 11.873671916727107
    slnshift 5.936835958363553
    slnmean 1.7811763245149566
    slnsd 0.2
  ]
  edge
  [
    source 2
    target 1
    length 102.75086968384085
    ttmedian 11.416763298204538
    slnshift 5.708381649102269
    slnmean 1.7419355595206691
    slnsd 0.2
  ]
  edge
  [
    source 3
    target 4
    length 100.1073433451236
    ttmedian 11.123038149458177
    slnshift 5.561519074729088
    slnmean 1.715871285825435
    slnsd 0.2
  ]
  edge
  [
    source 3
    target 13
    length 95.72888995788584
    ttmedian 10.636543328653982
    slnshift 5.318271664326991
    slnmean 1.6711483754436913
    slnsd 0.2
  ]
  edge
  [
    source 3
    target 2
    length 111.4333588202064
    ttmedian 12.381484313356266
    slnshift 6.190742156678133
    slnmean 1.823054975580326
    slnsd 0.2
  ]
  edge
  [
    source 4
    target 5
    length 135.08842032885718
    ttmedian 15.00982448098413
    slnshift 7.504912240492065
    slnmean 2.0155577715449797
    slnsd 0.2
  ]
  edge
  [
    source 4
    target 14
    length 78.56322922117194
    ttmedian 8.729247691241326
    slnshift 4.364623845620663
    slnmean 1.4735320104627359
    slnsd 0.2
  ]
  edge
  [
    source 4
    target 3
    length 129.31337285947671
    ttmedian 14.368152539941857
    slnshift 7.184076269970928
    slnmean 1.9718669475876025
    slnsd 0.2
  ]
  edge
  [
    source 5
    target 6
    length 105.88225962616991
    ttmedian 11.764695514018879
    slnshift 5.882347757009439
    slnmean 1.7719559606230917
    slnsd 0.2
  ]
  edge
  [
    source 5
    target 15
    length 111.771274274474
    ttmedian 12.419030474941556
    slnshift 6.209515237470778
    slnmean 1.8260828312995137
    slnsd 0.2
  ]
  edge
  [
    source 5
    target 4
    length 75.26533219920807
    ttmedian 8.362814688800897
    slnshift 4.181407344400449
    slnmean 1.4306478751341656
    slnsd 0.2
  ]
  edge
  [
    source 6
    target 7
    length 103.43613877444291
    ttmedian 11.492904308271434
    slnshift 5.746452154135717
    slnmean 1.7485826477040856
    slnsd 0.2
  ]
  edge
  [
    source 6
    target 16
    length 144.93776399396648
    ttmedian 16.104195999329608
    slnshift 8.052097999664804
    slnmean 2.085932678551072
    slnsd 0.2
  ]
  edge
  [
    source 6
    target 5
    length 76.34167317926838
    ttmedian 8.482408131029821
    slnshift 4.241204065514911
    slnmean 1.4448472066456077
    slnsd 0.2
  ]
  edge
  [
    source 7
    target 8
    length 60.99171991268532
    ttmedian 6.776857768076146
    slnshift 3.388428884038073
    slnmean 1.22036635792745
    slnsd 0.2
  ]
  edge
  [
    source 7
    target 17
    length 64.67192050471905
    ttmedian 7.185768944968784
    slnshift 3.592884472484392
    slnmean 1.2789553541172
    slnsd 0.2
  ]
  edge
  [
    source 7
    target 6
    length 106.87797113758141
    ttmedian 11.875330126397934
    slnshift 5.937665063198967
    slnmean 1.7813159690920641
    slnsd 0.2
  ]
  edge
  [
    source 8
    target 9
    length 181.52838245936752
    ttmedian 20.169820273263056
    slnshift 10.084910136631528
    slnmean 2.3110402607697007
    slnsd 0.2
  ]
  edge
  [
    source 8
    target 18
    length 132.43513777469084
    ttmedian 14.715015308298982
    slnshift 7.357507654149491
    slnmean 1.9957212414442824
    slnsd 0.2
  ]
  edge
  [
    source 8
    target 7
    length 151.9156965944018
    ttmedian 16.87952184382242
    slnshift 8.43976092191121
    slnmean 2.1329539814188836
    slnsd 0.2
  ]
  edge
  [
    source 9
    target 19
    length 102.50970632481038
    ttmedian 11.389967369423376
    slnshift 5.694983684711688
    slnmean 1.7395857320507377
    slnsd 0.2
  ]
  edge
  [
    source 9
    target 8
    length 141.8356907528674
    ttmedian 15.759521194763046
    slnshift 7.879760597381523
    slnmean 2.0642975223653206
    slnsd 0.2
  ]
  edge
  [
    source 10
    target 11
    length 93.44361338688341
    ttmedian 10.382623709653712
    slnshift 5.191311854826856
    slnmean 1.6469864311004814
    slnsd 0.2
  ]
  edge
  [
    source 10
    target 20
    length 116.56758484281038
    ttmedian 12.951953871423376
    slnshift 6.475976935711688
    slnmean 1.8680994743068364
    slnsd 0.2
  ]
  edge
  [
    source 10
    target 0
    length 93.53548554320662
    ttmedian 10.392831727022958
    slnshift 5.196415863511479
    slnmean 1.6479691309235345
    slnsd 0.2
  ]
  edge
  [
    source 11
    target 12
    length 87.16849174873946
    ttmedian 9.685387972082161
    slnshift 4.8426939860410805
    slnmean 1.5774711745655567
    slnsd 0.2
  ]
  edge
  [
    source 11
    target 21
    length 103.37335703120036
    ttmedian 11.485928559022263
    slnshift 5.742964279511131
    slnmean 1.7479755020319145
    slnsd 0.2
  ]
  edge
  [
    source 11
    target 1
    length 88.77762808314124
    ttmedian 9.864180898126804
    slnshift 4.932090449063402
    slnmean 1.5957629243600924
    slnsd 0.2
  ]
  edge
  [
    source 11
    target 10
    length 138.696550418994
    ttmedian 15.410727824332668
    slnshift 7.705363912166334
    slnmean 2.041916698307796
    slnsd 0.2
  ]
  edge
  [
    source 12
    target 13
    length 104.99613141334656
    ttmedian 11.666236823705173
    slnshift 5.8331184118525865
    slnmean 1.7635517479001999
    slnsd 0.2
  ]
  edge
  [
    source 12
    target 22
    length 110.52289378259756
    ttmedian 12.280321531399728
    slnshift 6.140160765699864
    slnmean 1.8148509251550136
    slnsd 0.2
  ]
  edge
  [
    source 12
    target 2
    length 91.90561272081997
    ttmedian 10.211734746757775
    slnshift 5.105867373378888
    slnmean 1.630390343820544
    slnsd 0.2
  ]
  edge
  [
    source 12
    target 11
    length 136.90513653404594
    ttmedian 15.211681837116215
    slnshift 7.6058409185581075
    slnmean 2.0289164940307156
    slnsd 0.2
  ]
  edge
  [
    source 13
    target 14
    length 83.27745052018486
    ttmedian 9.253050057798317
    slnshift 4.626525028899159
    slnmean 1.5318060525916246
    slnsd 0.2
  ]
  edge
  [
    source 13
    target 23
    length 117.94613931682365
    ttmedian 13.105126590758182
    slnshift 6.552563295379091
    slnmean 1.8798563158985049
    slnsd 0.2
  ]
  edge
  [
    source 13
    target 3
    length 91.60191825672584
    ttmedian 10.177990917413982
    slnshift 5.088995458706991
    slnmean 1.6270804552319589
    slnsd 0.2
  ]
  edge
  [
    source 13
    target 12
    length 79.06828498750174
    ttmedian 8.785364998611305
    slnshift 4.392682499305653
    slnmean 1.479940088151289
    slnsd 0.2
  ]
  edge
  [
    source 14
    target 15
    length 88.48525091229564
    ttmedian 9.831694545810628
    slnshift 4.915847272905314
    slnmean 1.592464123870764
    slnsd 0.2
  ]
  edge
  [
    source 14
    target 24
    length 81.77915648324047
    ttmedian 9.086572942582274
    slnshift 4.543286471291137
    slnmean 1.513650642524438
    slnsd 0.2
  ]
  edge
  [
    source 14
    target 4
    length 94.82146202534743
    ttmedian 10.535718002816381
    slnshift 5.2678590014081905
    slnmean 1.6616240184142987
    slnsd 0.2
  ]
  edge
  [
    source 14
    target 13
    length 91.87097292580208
    ttmedian 10.207885880644675
    slnshift 5.103942940322337
    slnmean 1.6300133665760679
    slnsd 0.2
  ]
  edge
  [
    source 15
    target 16
    length 108.11685901454433
    ttmedian 12.01298433494937
    slnshift 6.006492167474685
    slnmean 1.7928409121702322
    slnsd 0.2
  ]
  edge
  [
    source 15
    target 25
    length 115.17564988601583
    ttmedian 12.797294431779537
    slnshift 6.398647215889769
    slnmean 1.8560865955060837
    slnsd 0.2
  ]
  edge
  [
    source 15
    target 5
    length 96.38158164194675
    ttmedian 10.709064626882972
    slnshift 5.354532313441486
    slnmean 1.6779433636601715
    slnsd 0.2
  ]
  edge
  [
    source 15
    target 14
    length 99.35445920488087
    ttmedian 11.039384356097875
    slnshift 5.519692178048937
    slnmean 1.70832209388788
    slnsd 0.2
  ]
  edge
  [
    source 16
    target 17
    length 107.49493572671243
    ttmedian 11.943881747412492
    slnshift 5.971940873706246
    slnmean 1.7870719790429104
    slnsd 0.2
  ]
  edge
  [
    source 16
    target 26
    length 87.37423264125191
    ttmedian 9.708248071250212
    slnshift 4.854124035625106
    slnmean 1.5798286602587512
    slnsd 0.2
  ]
  edge
  [
    source 16
    target 6
    length 119.36788452722729
    ttmedian 13.263098280803032
    slnshift 6.631549140401516
    slnmean 1.8918384330735776
    slnsd 0.2
  ]
  edge
  [
    source 16
    target 15
    length 123.43644035388708
    ttmedian 13.715160039320786
    slnshift 6.857580019660393
    slnmean 1.9253546126875907
    slnsd 0.2
  ]
  edge
  [
    source 17
    target 18
    length 105.2212614919623
    ttmedian 11.6912512768847
    slnshift 5.84562563844235
    slnmean 1.7656936274237962
    slnsd 0.2
  ]
  edge
  [
    source 17
    target 27
    length 181.9707759006391
    ttmedian 20.21897510007101
    slnshift 10.109487550035505
    slnmean 2.3134743443130605
    slnsd 0.2
  ]
  edge
  [
    source 17
    target 7
    length 125.78007185188847
    ttmedian 13.975563539098719
    slnshift 6.987781769549359
    slnmean 1.9441631624695341
    slnsd 0.2
  ]
  edge
  [
    source 17
    target 16
    length 97.2322132102281
    ttmedian 10.803579245580899
    slnshift 5.401789622790449
    slnmean 1.6867303102933866
    slnsd 0.2
  ]
  edge
  [
    source 18
    target 19
    length 91.3433126372167
    ttmedian 10.149256959690744
    slnshift 5.074628479845372
    slnmean 1.6242533163063666
    slnsd 0.2
  ]
  edge
  [
    source 18
    target 28
    length 94.3606951963313
    ttmedian 10.484521688481257
    slnshift 5.242260844240628
    slnmean 1.656752864105904
    slnsd 0.2
  ]
  edge
  [
    source 18
    target 8
    length 88.2115547871696
    ttmedian 9.801283865241068
    slnshift 4.900641932620534
    slnmean 1.58936620319321
    slnsd 0.2
  ]
  edge
  [
    source 18
    target 17
    length 132.60389803502372
    ttmedian 14.733766448335968
    slnshift 7.366883224167984
    slnmean 1.996994716369396
    slnsd 0.2
  ]
  edge
  [
    source 19
    target 29
    length 83.99167429288616
    ttmedian 9.332408254765129
    slnshift 4.6662041273825645
    slnmean 1.540345920473998
    slnsd 0.2
  ]
  edge
  [
    source 19
    target 9
    length 97.99237211047783
    ttmedian 10.888041345608649
    slnshift 5.444020672804324
    slnmean 1.69451788213769
    slnsd 0.2
  ]
  edge
  [
    source 19
    target 18
    length 87.60570888493423
    ttmedian 9.733967653881582
    slnshift 4.866983826940791
    slnmean 1.5824744078420547
    slnsd 0.2
  ]
  edge
  [
    source 20
    target 21
    length 129.893632005209
    ttmedian 14.432625778356558
    slnshift 7.216312889178279
    slnmean 1.9763441422939927
    slnsd 0.2
  ]
  edge
  [
    source 20
    target 30
    length 70.14266073913646
    ttmedian 7.7936289710151625
    slnshift 3.8968144855075812
    slnmean 1.360159420786078
    slnsd 0.2
  ]
  edge
  [
    source 20
    target 10
    length 91.33603432637281
    ttmedian 10.148448258485868
    slnshift 5.074224129242934
    slnmean 1.6241736323030163
    slnsd 0.2
  ]
  edge
  [
    source 21
    target 22
    length 96.99833012971465
    ttmedian 10.77759223663496
    slnshift 5.38879611831748
    slnmean 1.6843220053014547
    slnsd 0.2
  ]
  edge
  [
    source 21
    target 31
    length 108.31127865185677
    ttmedian 12.034586516872974
    slnshift 6.017293258436487
    slnmean 1.7946375333647326
    slnsd 0.2
  ]
  edge
  [
    source 21
    target 11
    length 112.21267015054765
    ttmedian 12.46807446117196
    slnshift 6.23403723058598
    slnmean 1.830024153509856
    slnsd 0.2
  ]
  edge
  [
    source 21
    target 20
    length 94.74818734968117
    ttmedian 10.527576372186797
    slnshift 5.263788186093398
    slnmean 1.6608509550092962
    slnsd 0.2
  ]
  edge
  [
    source 22
    target 23
    length 128.05223932179547
    ttmedian 14.228026591310607
    slnshift 7.114013295655304
    slnmean 1.9620665424667862
    slnsd 0.2
  ]
  edge
  [
    source 22
    target 32
    length 108.18672080988141
    ttmedian 12.02074675665349
    slnshift 6.010373378326745
    slnmean 1.79348687279486
    slnsd 0.2
  ]
  edge
  [
    source 22
    target 12
    length 185.3197942112042
    ttmedian 20.591088245689356
    slnshift 10.295544122844678
    slnmean 2.3317111922005775
    slnsd 0.2
  ]
  edge
  [
    source 22
    target 21
    length 68.58609478081408
    ttmedian 7.620677197868231
    slnshift 3.8103385989341154
    slnmean 1.3377180563011226
    slnsd 0.2
  ]
  edge
  [
    source 23
    target 24
    length 116.78333165069705
    ttmedian 12.975925738966339
    slnshift 6.4879628694831695
    slnmean 1.8699485938440799
    slnsd 0.2
  ]
  edge
  [
    source 23
    target 33
    length 77.006285747745
    ttmedian 8.556253971971666
    slnshift 4.278126985985833
    slnmean 1.4535152937133173
    slnsd 0.2
  ]
  edge
  [
    source 23
    target 13
    length 81.9138415087417
    ttmedian 9.101537945415744
    slnshift 4.550768972707872
    slnmean 1.5152962236738656
    slnsd 0.2
  ]
  edge
  [
    source 23
    target 22
    length 164.25848808112303
    ttmedian 18.25094312012478
    slnshift 9.12547156006239
    slnmean 2.2110695759483594
    slnsd 0.2
  ]
  edge
  [
    source 24
    target 25
    length 154.68235986771674
    ttmedian 17.186928874190748
    slnshift 8.593464437095374
    slnmean 2.151001965176752
    slnsd 0.2
  ]
  edge
  [
    source 24
    target 34
    length 62.87343992809813
    ttmedian 6.985937769788681
    slnshift 3.4929688848943403
    slnmean 1.2507520579391396
    slnsd 0.2
  ]
  edge
  [
    source 24
    target 14
    length 94.58290592024005
    ttmedian 10.509211768915561
    slnshift 5.2546058844577805
    slnmean 1.6591050033161767
    slnsd 0.2
  ]
  edge
  [
    source 24
    target 23
    length 98.36727956958694
    ttmedian 10.929697729954105
    slnshift 5.464848864977053
    slnmean 1.6983364661641265
    slnsd 0.2
  ]
  edge
  [
    source 25
    target 26
    length 58.68840390592073
    ttmedian 6.520933767324525
    slnshift 3.2604668836622626
    slnmean 1.1818704009715257
    slnsd 0.2
  ]
  edge
  [
    source 25
    target 35
    length 98.7866582967747
    ttmedian 10.9762953663083
    slnshift 5.48814768315415
    slnmean 1.7025908002572383
    slnsd 0.2
  ]
  edge
  [
    source 25
    target 15
    length 110.3332675249892
    ttmedian 12.259251947221022
    slnshift 6.129625973610511
    slnmean 1.8131337323623617
    slnsd 0.2
  ]
  edge
  [
    source 25
    target 24
    length 105.58184902523043
    ttmedian 11.731316558358936
    slnshift 5.865658279179468
    slnmean 1.7691147143714365
    slnsd 0.2
  ]
  edge
  [
    source 26
    target 27
    length 60.738134334878616
    ttmedian 6.7486815927642905
    slnshift 3.3743407963821452
    slnmean 1.2161999856566623
    slnsd 0.2
  ]
  edge
  [
    source 26
    target 36
    length 131.9070152828401
    ttmedian 14.656335031426677
    slnshift 7.328167515713338
    slnmean 1.9917254867875893
    slnsd 0.2
  ]
  edge
  [
    source 26
    target 16
    length 106.3043472063609
    ttmedian 11.8115941340401
    slnshift 5.90579706702005
    slnmean 1.7759344222539242
    slnsd 0.2
  ]
  edge
  [
    source 26
    target 25
    length 98.46411613593024
    ttmedian 10.940457348436693
    slnshift 5.470228674218347
    slnmean 1.6993204207188257
    slnsd 0.2
  ]
  edge
  [
    source 27
    target 28
    length 82.82226662258375
    ttmedian 9.202474069175972
    slnshift 4.601237034587986
    slnmean 1.5263251879050017
    slnsd 0.2
  ]
  edge
  [
    source 27
    target 37
    length 119.48130792725537
    ttmedian 13.275700880806152
    slnshift 6.637850440403076
    slnmean 1.8927881822218866
    slnsd 0.2
  ]
  edge
  [
    source 27
    target 17
    length 125.80555162511848
    ttmedian 13.978394625013165
    slnshift 6.989197312506582
    slnmean 1.9443657159617804
    slnsd 0.2
  ]
  edge
  [
    source 27
    target 26
    length 88.6456446224792
    ttmedian 9.849516069164356
    slnshift 4.924758034582178
    slnmean 1.5942751433830569
    slnsd 0.2
  ]
  edge
  [
    source 28
    target 29
    length 102.26495967746463
    ttmedian 11.36277329749607
    slnshift 5.681386648748035
    slnmean 1.7371953312393058
    slnsd 0.2
  ]
  edge
  [
    source 28
    target 38
    length 122.96746911521488
    ttmedian 13.663052123912765
    slnshift 6.8315260619563825
    slnmean 1.921548083757493
    slnsd 0.2
  ]
  edge
  [
    source 28
    target 18
    length 61.34674760558313
    ttmedian 6.816305289509237
    slnshift 3.4081526447546184
    slnmean 1.2261703980915504
    slnsd 0.2
  ]
  edge
  [
    source 28
    target 27
    length 102.98044349737827
    ttmedian 11.442271499708697
    slnshift 5.721135749854349
    slnmean 1.7441673433485863
    slnsd 0.2
  ]
  edge
  [
    source 29
    target 39
    length 62.07268045967129
    ttmedian 6.896964495519033
    slnshift 3.4484822477595163
    slnmean 1.2379342060608265
    slnsd 0.2
  ]
  edge
  [
    source 29
    target 19
    length 79.39559354318449
    ttmedian 8.821732615909388
    slnshift 4.410866307954694
    slnmean 1.4840711118796903
    slnsd 0.2
  ]
  edge
  [
    source 29
    target 28
    length 112.46117427350448
    ttmedian 12.495686030389386
    slnshift 6.247843015194693
    slnmean 1.8322362866126873
    slnsd 0.2
  ]
  edge
  [
    source 30
    target 31
    length 96.45518489394664
    ttmedian 10.71724276599407
    slnshift 5.358621382997035
    slnmean 1.6787067373351858
    slnsd 0.2
  ]
  edge
  [
    source 30
    target 40
    length 90.48061220638085
    ttmedian 10.053401356264539
    slnshift 5.026700678132269
    slnmean 1.6147638400954945
    slnsd 0.2
  ]
  edge
  [
    source 30
    target 20
    length 92.86754426784269
    ttmedian 10.3186160297603
    slnshift 5.15930801488015
    slnmean 1.6408024648571238
    slnsd 0.2
  ]
  edge
  [
    source 31
    target 32
    length 123.6212354659838
    ttmedian 13.735692829553756
    slnshift 6.867846414776878
    slnmean 1.9268505803473406
    slnsd 0.2
  ]
  edge
  [
    source 31
    target 41
    length 119.32821427504648
    ttmedian 13.258690475005164
    slnshift 6.629345237502582
    slnmean 1.8915060417798353
    slnsd 0.2
  ]
  edge
  [
    source 31
    target 21
    length 82.13407347586613
    ttmedian 9.126008163985126
    slnshift 4.563004081992563
    slnmean 1.5179811965301302
    slnsd 0.2
  ]
  edge
  [
    source 31
    target 30
    length 107.59835582736169
    ttmedian 11.955372869706855
    slnshift 5.977686434853427
    slnmean 1.7880336092999185
    slnsd 0.2
  ]
  edge
  [
    source 32
    target 33
    length 88.89610898521457
    ttmedian 9.87734544280162
    slnshift 4.93867272140081
    slnmean 1.5970966152205475
    slnsd 0.2
  ]
  edge
  [
    source 32
    target 42
    length 182.40903081694816
    ttmedian 20.267670090772018
    slnshift 10.133835045386009
    slnmean 2.3158798295762644
    slnsd 0.2
  ]
  edge
  [
    source 32
    target 22
    length 83.12379489448084
    ttmedian 9.235977210497872
    slnshift 4.617988605248936
    slnmean 1.5299592434647586
    slnsd 0.2
  ]
  edge
  [
    source 32
    target 31
    length 92.47728036315753
    ttmedian 10.275253373684171
    slnshift 5.1376266868420855
    slnmean 1.6365912387576542
    slnsd 0.2
  ]
  edge
  [
    source 33
    target 34
    length 91.64750093543678
    ttmedian 10.183055659492975
    slnshift 5.091527829746488
    slnmean 1.6275779485375532
    slnsd 0.2
  ]
  edge
  [
    source 33
    target 43
    length 89.60522445522695
    ttmedian 9.956136050580772
    slnshift 4.978068025290386
    slnmean 1.6050418690369233
    slnsd 0.2
  ]
  edge
  [
    source 33
    target 23
    length 103.59056228122863
    ttmedian 11.51006247569207
    slnshift 5.755031237846035
    slnmean 1.7500744701080349
    slnsd 0.2
  ]
  edge
  [
    source 33
    target 32
    length 107.06461342685596
    ttmedian 11.896068158539551
    slnshift 5.9480340792697755
    slnmean 1.7830607581157545
    slnsd 0.2
  ]
  edge
  [
    source 34
    target 35
    length 66.72151729568155
    ttmedian 7.4135019217423945
    slnshift 3.7067509608711973
    slnmean 1.310155741137978
    slnsd 0.2
  ]
  edge
  [
    source 34
    target 44
    length 86.6512849802073
    ttmedian 9.627920553356367
    slnshift 4.813960276678183
    slnmean 1.5715200877020292
    slnsd 0.2
  ]
  edge
  [
    source 34
    target 24
    length 71.88433000909195
    ttmedian 7.987147778787995
    slnshift 3.9935738893939976
    slnmean 1.3846865416190626
    slnsd 0.2
  ]
  edge
  [
    source 34
    target 33
    length 136.2132289066106
    ttmedian 15.134803211845622
    slnshift 7.567401605922811
    slnmean 2.023849759638631
    slnsd 0.2
  ]
  edge
  [
    source 35
    target 36
    length 185.21206326362565
    ttmedian 20.57911814040285
    slnshift 10.289559070201426
    slnmean 2.331129698607296
    slnsd 0.2
  ]
  edge
  [
    source 35
    target 45
    length 141.31313496565667
    ttmedian 15.701459440628518
    slnshift 7.850729720314259
    slnmean 2.060606485475806
    slnsd 0.2
  ]
  edge
  [
    source 35
    target 25
    length 109.04149139799621
    ttmedian 12.115721266444023
    slnshift 6.0578606332220115
    slnmean 1.8013567069297223
    slnsd 0.2
  ]
  edge
  [
    source 35
    target 34
    length 129.12736111828661
    ttmedian 14.347484568698512
    slnshift 7.173742284349256
    slnmean 1.9704274549004606
    slnsd 0.2
  ]
  edge
  [
    source 36
    target 37
    length 104.25843333877434
    ttmedian 11.584270370974927
    slnshift 5.792135185487464
    slnmean 1.75650099484686
    slnsd 0.2
  ]
  edge
  [
    source 36
    target 46
    length 151.31281887771888
    ttmedian 16.812535430857654
    slnshift 8.406267715428827
    slnmean 2.1289775842114995
    slnsd 0.2
  ]
  edge
  [
    source 36
    target 26
    length 118.1721651950409
    ttmedian 13.130240577226767
    slnshift 6.565120288613383
    slnmean 1.8817708303002874
    slnsd 0.2
  ]
  edge
  [
    source 36
    target 35
    length 94.41344841654848
    ttmedian 10.490383157394275
    slnshift 5.2451915786971375
    slnmean 1.657311767148246
    slnsd 0.2
  ]
  edge
  [
    source 37
    target 38
    length 75.39812349572725
    ttmedian 8.377569277303028
    slnshift 4.188784638651514
    slnmean 1.4324106294834789
    slnsd 0.2
  ]
  edge
  [
    source 37
    target 47
    length 85.21517556470562
    ttmedian 9.468352840522847
    slnshift 4.734176420261424
    slnmean 1.5548077770348023
    slnsd 0.2
  ]
  edge
  [
    source 37
    target 27
    length 108.16175264783769
    ttmedian 12.01797251642641
    slnshift 6.008986258213205
    slnmean 1.7932560584816781
    slnsd 0.2
  ]
  edge
  [
    source 37
    target 36
    length 73.60486991016698
    ttmedian 8.178318878907442
    slnshift 4.089159439453721
    slnmean 1.4083394329074128
    slnsd 0.2
  ]
  edge
  [
    source 38
    target 39
    length 94.6060948798862
    ttmedian 10.511788319987355
    slnshift 5.2558941599936775
    slnmean 1.6593501439920533
    slnsd 0.2
  ]
  edge
  [
    source 38
    target 48
    length 139.7914631888107
    ttmedian 15.532384798756745
    slnshift 7.766192399378372
    slnmean 2.049780005581315
    slnsd 0.2
  ]
  edge
  [
    source 38
    target 28
    length 100.73543390244227
    ttmedian 11.192825989160252
    slnshift 5.596412994580126
    slnmean 1.722125855827719
    slnsd 0.2
  ]
  edge
  [
    source 38
    target 37
    length 164.27085534910748
    ttmedian 18.252317261011942
    slnshift 9.126158630505971
    slnmean 2.211144864614274
    slnsd 0.2
  ]
  edge
  [
    source 39
    target 49
    length 143.5899382346113
    ttmedian 15.954437581623479
    slnshift 7.977218790811739
    slnmean 2.0765898282589825
    slnsd 0.2
  ]
  edge
  [
    source 39
    target 29
    length 93.0617110316809
    ttmedian 10.340190114631211
    slnshift 5.1700950573156055
    slnmean 1.6428910746792815
    slnsd 0.2
  ]
  edge
  [
    source 39
    target 38
    length 71.18930117361568
    ttmedian 7.909922352623964
    slnshift 3.954961176311982
    slnmean 1.374970784815427
    slnsd 0.2
  ]
  edge
  [
    source 40
    target 41
    length 98.80615048421072
    ttmedian 10.978461164912302
    slnshift 5.489230582456151
    slnmean 1.7027880967845714
    slnsd 0.2
  ]
  edge
  [
    source 40
    target 50
    length 88.67458768840315
    ttmedian 9.852731965378128
    slnshift 4.926365982689064
    slnmean 1.5946015930589208
    slnsd 0.2
  ]
  edge
  [
    source 40
    target 30
    length 109.90413188574747
    ttmedian 12.211570209527498
    slnshift 6.105785104763749
    slnmean 1.809236699583262
    slnsd 0.2
  ]
  edge
  [
    source 41
    target 42
    length 149.72079299231774
    ttmedian 16.63564366581308
    slnshift 8.31782183290654
    slnmean 2.118400421628293
    slnsd 0.2
  ]
  edge
  [
    source 41
    target 51
    length 75.51974741205795
    ttmedian 8.391083045784217
    slnshift 4.195541522892109
    slnmean 1.4340224192858309
    slnsd 0.2
  ]
  edge
  [
    source 41
    target 31
    length 90.73908580260517
    ttmedian 10.082120644733909
    slnshift 5.041060322366954
    slnmean 1.617616441378224
    slnsd 0.2
  ]
  edge
  [
    source 41
    target 40
    length 108.66368714057964
    ttmedian 12.07374301561996
    slnshift 6.03687150780998
    slnmean 1.797885915474806
    slnsd 0.2
  ]
  edge
  [
    source 42
    target 43
    length 132.80919568467704
    ttmedian 14.756577298297449
    slnshift 7.3782886491487245
    slnmean 1.998541721362403
    slnsd 0.2
  ]
  edge
  [
    source 42
    target 52
    length 113.61877764950366
    ttmedian 12.624308627722629
    slnshift 6.312154313861314
    slnmean 1.8424770309413925
    slnsd 0.2
  ]
  edge
  [
    source 42
    target 32
    length 110.91283208640148
    ttmedian 12.323648009600165
    slnshift 6.161824004800082
    slnmean 1.8183728383904292
    slnsd 0.2
  ]
  edge
  [
    source 42
    target 41
    length 139.95312454012148
    ttmedian 15.550347171124608
    slnshift 7.775173585562304
    slnmean 2.050935783933412
    slnsd 0.2
  ]
  edge
  [
    source 43
    target 44
    length 113.13129142725013
    ttmedian 12.570143491916681
    slnshift 6.285071745958341
    slnmean 1.838177257403999
    slnsd 0.2
  ]
  edge
  [
    source 43
    target 53
    length 94.2470135974323
    ttmedian 10.4718903997147
    slnshift 5.23594519985735
    slnmean 1.6555473819610564
    slnsd 0.2
  ]
  edge
  [
    source 43
    target 33
    length 101.44243284827819
    ttmedian 11.271381427586466
    slnshift 5.635690713793233
    slnmean 1.7291197156309313
    slnsd 0.2
  ]
  edge
  [
    source 43
    target 42
    length 83.79742878169434
    ttmedian 9.310825420188259
    slnshift 4.655412710094129
    slnmean 1.5380305663251357
    slnsd 0.2
  ]
  edge
  [
    source 44
    target 45
    length 114.64454119263327
    ttmedian 12.73828235473703
    slnshift 6.369141177368515
    slnmean 1.8514646374755603
    slnsd 0.2
  ]
  edge
  [
    source 44
    target 54
    length 126.56997132161689
    ttmedian 14.06333014684632
    slnshift 7.03166507342316
    slnmean 1.9504235303269617
    slnsd 0.2
  ]
  edge
  [
    source 44
    target 34
    length 47.47821590041439
    ttmedian 5.275357322268266
    slnshift 2.637678661134133
    slnmean 0.9698992353266433
    slnsd 0.2
  ]
  edge
  [
    source 44
    target 43
    length 135.59973192849702
    ttmedian 15.066636880944113
    slnshift 7.533318440472057
    slnmean 2.0193356406796883
    slnsd 0.2
  ]
  edge
  [
    source 45
    target 46
    length 98.59691816799673
    ttmedian 10.955213129777414
    slnshift 5.477606564888707
    slnmean 1.700668247321279
    slnsd 0.2
  ]
  edge
  [
    source 45
    target 55
    length 112.21226480340766
    ttmedian 12.468029422600852
    slnshift 6.234014711300426
    slnmean 1.8300205411916433
    slnsd 0.2
  ]
  edge
  [
    source 45
    target 35
    length 118.00932997435282
    ttmedian 13.112147774928092
    slnshift 6.556073887464046
    slnmean 1.880391931023085
    slnsd 0.2
  ]
  edge
  [
    source 45
    target 44
    length 56.33941357684938
    ttmedian 6.259934841872154
    slnshift 3.129967420936077
    slnmean 1.1410225958512707
    slnsd 0.2
  ]
  edge
  [
    source 46
    target 47
    length 74.15725144881492
    ttmedian 8.23969460542388
    slnshift 4.11984730271194
    slnmean 1.4158161002260472
    slnsd 0.2
  ]
  edge
  [
    source 46
    target 56
    length 92.00978181320666
    ttmedian 10.223309090356295
    slnshift 5.111654545178148
    slnmean 1.631523137557468
    slnsd 0.2
  ]
  edge
  [
    source 46
    target 36
    length 82.04948109365706
    ttmedian 9.116609010406341
    slnshift 4.5583045052031705
    slnmean 1.5169507353500657
    slnsd 0.2
  ]
  edge
  [
    source 46
    target 45
    length 107.09468837060444
    ttmedian 11.89940981895605
    slnshift 5.949704909478025
    slnmean 1.7833416232819534
    slnsd 0.2
  ]
  edge
  [
    source 47
    target 48
    length 87.92135356298951
    ttmedian 9.769039284776612
    slnshift 4.884519642388306
    slnmean 1.5860709474741137
    slnsd 0.2
  ]
  edge
  [
    source 47
    target 57
    length 65.19627834462426
    ttmedian 7.244030927180473
    slnshift 3.6220154635902366
    slnmean 1.2870306288033606
    slnsd 0.2
  ]
  edge
  [
    source 47
    target 37
    length 116.54115985821139
    ttmedian 12.949017762023487
    slnshift 6.474508881011744
    slnmean 1.8678727562165642
    slnsd 0.2
  ]
  edge
  [
    source 47
    target 46
    length 131.65732567705112
    ttmedian 14.628591741894569
    slnshift 7.3142958709472845
    slnmean 1.9898307715952588
    slnsd 0.2
  ]
  edge
  [
    source 48
    target 49
    length 115.15261293490462
    ttmedian 12.794734770544958
    slnshift 6.397367385272479
    slnmean 1.8558865596884562
    slnsd 0.2
  ]
  edge
  [
    source 48
    target 58
    length 83.67751695540777
    ttmedian 9.297501883934196
    slnshift 4.648750941967098
    slnmean 1.5365985688851616
    slnsd 0.2
  ]
  edge
  [
    source 48
    target 38
    length 93.6866137660423
    ttmedian 10.409623751782478
    slnshift 5.204811875891239
    slnmean 1.649583558453383
    slnsd 0.2
  ]
  edge
  [
    source 48
    target 47
    length 114.69508847070553
    ttmedian 12.74389871896728
    slnshift 6.37194935948364
    slnmean 1.8519054446658827
    slnsd 0.2
  ]
  edge
  [
    source 49
    target 59
    length 116.27638059191807
    ttmedian 12.919597843546452
    slnshift 6.459798921773226
    slnmean 1.8655981906495747
    slnsd 0.2
  ]
  edge
  [
    source 49
    target 39
    length 128.6229775794368
    ttmedian 14.291441953270756
    slnshift 7.145720976635378
    slnmean 1.9665137127481325
    slnsd 0.2
  ]
  edge
  [
    source 49
    target 48
    length 92.90371963446698
    ttmedian 10.322635514940776
    slnshift 5.161317757470388
    slnmean 1.6411919262444412
    slnsd 0.2
  ]
  edge
  [
    source 50
    target 51
    length 70.10191117349727
    ttmedian 7.789101241499697
    slnshift 3.8945506207498486
    slnmean 1.359578299303368
    slnsd 0.2
  ]
  edge
  [
    source 50
    target 60
    length 84.38869881439436
    ttmedian 9.376522090488262
    slnshift 4.688261045244131
    slnmean 1.5450617344369164
    slnsd 0.2
  ]
  edge
  [
    source 50
    target 40
    length 114.27883560166808
    ttmedian 12.697648400185342
    slnshift 6.348824200092671
    slnmean 1.8482696304196453
    slnsd 0.2
  ]
  edge
  [
    source 51
    target 52
    length 120.44145532397854
    ttmedian 13.382383924886504
    slnshift 6.691191962443252
    slnmean 1.9007920290384077
    slnsd 0.2
  ]
  edge
  [
    source 51
    target 61
    length 174.4141416611205
    ttmedian 19.379349073457835
    slnshift 9.689674536728917
    slnmean 2.2710608377977253
    slnsd 0.2
  ]
  edge
  [
    source 51
    target 41
    length 102.97286851314283
    ttmedian 11.441429834793647
    slnshift 5.7207149173968235
    slnmean 1.7440937831405476
    slnsd 0.2
  ]
  edge
  [
    source 51
    target 50
    length 106.30622199618136
    ttmedian 11.811802444020152
    slnshift 5.905901222010076
    slnmean 1.7759520581581743
    slnsd 0.2
  ]
  edge
  [
    source 52
    target 53
    length 95.66432795475342
    ttmedian 10.62936977275038
    slnshift 5.31468488637519
    slnmean 1.6704737224305348
    slnsd 0.2
  ]
  edge
  [
    source 52
    target 62
    length 90.35422730773627
    ttmedian 10.039358589748474
    slnshift 5.019679294874237
    slnmean 1.6133660461796537
    slnsd 0.2
  ]
  edge
  [
    source 52
    target 42
    length 121.58504582806668
    ttmedian 13.509449536451854
    slnshift 6.754724768225927
    slnmean 1.9102422256858305
    slnsd 0.2
  ]
  edge
  [
    source 52
    target 51
    length 109.2391580723814
    ttmedian 12.1376842302646
    slnshift 6.0688421151323
    slnmean 1.8031678315428892
    slnsd 0.2
  ]
  edge
  [
    source 53
    target 54
    length 94.94998834958226
    ttmedian 10.54999870550914
    slnshift 5.27499935275457
    slnmean 1.6629785566615671
    slnsd 0.2
  ]
  edge
  [
    source 53
    target 63
    length 76.34372492416388
    ttmedian 8.482636102684875
    slnshift 4.241318051342438
    slnmean 1.4448740821022144
    slnsd 0.2
  ]
  edge
  [
    source 53
    target 43
    length 96.97010439253772
    ttmedian 10.774456043615302
    slnshift 5.387228021807651
    slnmean 1.6840309709643917
    slnsd 0.2
  ]
  edge
  [
    source 53
    target 52
    length 90.68796332517036
    ttmedian 10.076440369463374
    slnshift 5.038220184731687
    slnmean 1.6170528817653258
    slnsd 0.2
  ]
  edge
  [
    source 54
    target 55
    length 136.86069982049963
    ttmedian 15.20674442449996
    slnshift 7.60337221224998
    slnmean 2.0285918610193194
    slnsd 0.2
  ]
  edge
  [
    source 54
    target 64
    length 126.71642369092919
    ttmedian 14.079602632325466
    slnshift 7.039801316162733
    slnmean 1.9515799476357236
    slnsd 0.2
  ]
  edge
  [
    source 54
    target 44
    length 77.44689788415316
    ttmedian 8.605210876017019
    slnshift 4.302605438008509
    slnmean 1.4592207550225782
    slnsd 0.2
  ]
  edge
  [
    source 54
    target 53
    length 133.88246364842144
    ttmedian 14.875829294269048
    slnshift 7.437914647134524
    slnmean 2.006590520192956
    slnsd 0.2
  ]
  edge
  [
    source 55
    target 56
    length 86.67590932043443
    ttmedian 9.630656591159381
    slnshift 4.8153282955796906
    slnmean 1.571804224769985
    slnsd 0.2
  ]
  edge
  [
    source 55
    target 65
    length 103.38444961623262
    ttmedian 11.487161068470291
    slnshift 5.743580534235146
    slnmean 1.7480828023096913
    slnsd 0.2
  ]
  edge
  [
    source 55
    target 45
    length 135.07850858498253
    ttmedian 15.00872317610917
    slnshift 7.504361588054585
    slnmean 2.015484396584403
    slnsd 0.2
  ]
  edge
  [
    source 55
    target 54
    length 77.3996256659515
    ttmedian 8.599958407327945
    slnshift 4.299979203663972
    slnmean 1.4586101863306058
    slnsd 0.2
  ]
  edge
  [
    source 56
    target 57
    length 104.09146426476752
    ttmedian 11.565718251640837
    slnshift 5.782859125820418
    slnmean 1.7548982188275664
    slnsd 0.2
  ]
  edge
  [
    source 56
    target 66
    length 75.380706793652
    ttmedian 8.375634088183554
    slnshift 4.187817044091777
    slnmean 1.4321796063070913
    slnsd 0.2
  ]
  edge
  [
    source 56
    target 46
    length 61.54971767707694
    ttmedian 6.838857519675216
    slnshift 3.419428759837608
    slnmean 1.2294735078363197
    slnsd 0.2
  ]
  edge
  [
    source 56
    target 55
    length 127.97304300449144
    ttmedian 14.21922700049905
    slnshift 7.109613500249525
    slnmean 1.961447882316426
    slnsd 0.2
  ]
  edge
  [
    source 57
    target 58
    length 100.70818000071573
    ttmedian 11.189797777857303
    slnshift 5.594898888928651
    slnmean 1.7218552699167065
    slnsd 0.2
  ]
  edge
  [
    source 57
    target 67
    length 81.4165269782284
    ttmedian 9.046280775358712
    slnshift 4.523140387679356
    slnmean 1.5092065286342604
    slnsd 0.2
  ]
  edge
  [
    source 57
    target 47
    length 67.257749482001
    ttmedian 7.473083275777889
    slnshift 3.7365416378889447
    slnmean 1.3181604878799646
    slnsd 0.2
  ]
  edge
  [
    source 57
    target 56
    length 90.45043138456256
    ttmedian 10.050047931618062
    slnshift 5.025023965809031
    slnmean 1.6144302232489964
    slnsd 0.2
  ]
  edge
  [
    source 58
    target 59
    length 139.8702220103212
    ttmedian 15.541135778924579
    slnshift 7.770567889462289
    slnmean 2.050343249155683
    slnsd 0.2
  ]
  edge
  [
    source 58
    target 68
    length 110.5853321247994
    ttmedian 12.28725912497771
    slnshift 6.143629562488855
    slnmean 1.8154157014640384
    slnsd 0.2
  ]
  edge
  [
    source 58
    target 48
    length 91.10200673764257
    ttmedian 10.122445193071396
    slnshift 5.061222596535698
    slnmean 1.62160807398271
    slnsd 0.2
  ]
  edge
  [
    source 58
    target 57
    length 84.5348881672072
    ttmedian 9.392765351911912
    slnshift 4.696382675955956
    slnmean 1.5467925689748403
    slnsd 0.2
  ]
  edge
  [
    source 59
    target 69
    length 102.68060074999981
    ttmedian 11.408955638888868
    slnshift 5.704477819444434
    slnmean 1.7412514487908655
    slnsd 0.2
  ]
  edge
  [
    source 59
    target 49
    length 67.92090308768535
    ttmedian 7.546767009742816
    slnshift 3.773383504871408
    slnmean 1.3279720803625443
    slnsd 0.2
  ]
  edge
  [
    source 59
    target 58
    length 139.9621198706007
    ttmedian 15.551346652288967
    slnshift 7.775673326144483
    slnmean 2.0510000557489843
    slnsd 0.2
  ]
  edge
  [
    source 60
    target 61
    length 113.40522673819345
    ttmedian 12.600580748688161
    slnshift 6.300290374344081
    slnmean 1.8405957235010488
    slnsd 0.2
  ]
  edge
  [
    source 60
    target 70
    length 134.7953680989247
    ttmedian 14.977263122102746
    slnshift 7.488631561051373
    slnmean 2.0133860787077382
    slnsd 0.2
  ]
  edge
  [
    source 60
    target 50
    length 86.85966481506222
    ttmedian 9.651073868340246
    slnshift 4.825536934170123
    slnmean 1.5739220102928908
    slnsd 0.2
  ]
  edge
  [
    source 61
    target 62
    length 123.82600256455774
    ttmedian 13.758444729395304
    slnshift 6.879222364697652
    slnmean 1.9285056171700936
    slnsd 0.2
  ]
  edge
  [
    source 61
    target 71
    length 118.76495447458323
    ttmedian 13.196106052731471
    slnshift 6.5980530263657355
    slnmean 1.8867746095073077
    slnsd 0.2
  ]
  edge
  [
    source 61
    target 51
    length 68.16307571404171
    ttmedian 7.573675079337968
    slnshift 3.786837539668984
    slnmean 1.331531248519381
    slnsd 0.2
  ]
  edge
  [
    source 61
    target 60
    length 103.79422789444794
    ttmedian 11.532691988271994
    slnshift 5.766345994135997
    slnmean 1.7520386033361486
    slnsd 0.2
  ]
  edge
  [
    source 62
    target 63
    length 108.19638743879703
    ttmedian 12.021820826533004
    slnshift 6.010910413266502
    slnmean 1.7935762201473433
    slnsd 0.2
  ]
  edge
  [
    source 62
    target 72
    length 86.61220751194182
    ttmedian 9.62357861243798
    slnshift 4.81178930621899
    slnmean 1.571069012078815
    slnsd 0.2
  ]
  edge
  [
    source 62
    target 52
    length 91.38920848346754
    ttmedian 10.154356498163061
    slnshift 5.0771782490815305
    slnmean 1.6247556444845637
    slnsd 0.2
  ]
  edge
  [
    source 62
    target 61
    length 111.52140104672337
    ttmedian 12.391266782969263
    slnshift 6.195633391484631
    slnmean 1.8238447522251113
    slnsd 0.2
  ]
  edge
  [
    source 63
    target 64
    length 99.0565318041629
    ttmedian 11.006281311573655
    slnshift 5.503140655786828
    slnmean 1.7053189575886682
    slnsd 0.2
  ]
  edge
  [
    source 63
    target 73
    length 77.78830494619905
    ttmedian 8.643144994022116
    slnshift 4.321572497011058
    slnmean 1.4636193399604027
    slnsd 0.2
  ]
  edge
  [
    source 63
    target 53
    length 115.93543689244842
    ttmedian 12.881715210272047
    slnshift 6.440857605136023
    slnmean 1.8626616997495073
    slnsd 0.2
  ]
  edge
  [
    source 63
    target 62
    length 105.86074742923837
    ttmedian 11.762305269915375
    slnshift 5.8811526349576875
    slnmean 1.7717527690532286
    slnsd 0.2
  ]
  edge
  [
    source 64
    target 65
    length 106.42759028250748
    ttmedian 11.825287809167499
    slnshift 5.912643904583749
    slnmean 1.7770930925635107
    slnsd 0.2
  ]
  edge
  [
    source 64
    target 74
    length 102.42296911056168
    ttmedian 11.38032990117352
    slnshift 5.69016495058676
    slnmean 1.7387392372763562
    slnsd 0.2
  ]
  edge
  [
    source 64
    target 54
    length 111.87986018676031
    ttmedian 12.431095576306701
    slnshift 6.2155477881533505
    slnmean 1.8270538607653826
    slnsd 0.2
  ]
  edge
  [
    source 64
    target 63
    length 93.09958103140629
    ttmedian 10.344397892378476
    slnshift 5.172198946189238
    slnmean 1.643297926177013
    slnsd 0.2
  ]
  edge
  [
    source 65
    target 66
    length 80.59854612018911
    ttmedian 8.955394013354345
    slnshift 4.477697006677173
    slnmean 1.4991088532426742
    slnsd 0.2
  ]
  edge
  [
    source 65
    target 75
    length 83.07252124968572
    ttmedian 9.230280138853969
    slnshift 4.615140069426984
    slnmean 1.5293422183993248
    slnsd 0.2
  ]
  edge
  [
    source 65
    target 55
    length 131.83669136726937
    ttmedian 14.648521263029929
    slnshift 7.3242606315149645
    slnmean 1.991192212131443
    slnsd 0.2
  ]
  edge
  [
    source 65
    target 64
    length 60.38498750587069
    ttmedian 6.709443056207855
    slnshift 3.3547215281039273
    slnmean 1.2103687649233226
    slnsd 0.2
  ]
  edge
  [
    source 66
    target 67
    length 114.46915113174853
    ttmedian 12.718794570194282
    slnshift 6.359397285097141
    slnmean 1.8499336063672427
    slnsd 0.2
  ]
  edge
  [
    source 66
    target 76
    length 69.72754722773828
    ttmedian 7.747505247526476
    slnshift 3.873752623763238
    slnmean 1.354223707438272
    slnsd 0.2
  ]
  edge
  [
    source 66
    target 56
    length 66.88385417151166
    ttmedian 7.4315393523901845
    slnshift 3.7157696761950922
    slnmean 1.3125858373823975
    slnsd 0.2
  ]
  edge
  [
    source 66
    target 65
    length 77.75229322961886
    ttmedian 8.639143692179873
    slnshift 4.319571846089937
    slnmean 1.4631562876429245
    slnsd 0.2
  ]
  edge
  [
    source 67
    target 68
    length 93.76454315326393
    ttmedian 10.41828257258488
    slnshift 5.20914128629244
    slnmean 1.65041502189156
    slnsd 0.2
  ]
  edge
  [
    source 67
    target 77
    length 120.0366421921488
    ttmedian 13.337404688016534
    slnshift 6.668702344008267
    slnmean 1.8974252898768094
    slnsd 0.2
  ]
  edge
  [
    source 67
    target 57
    length 65.40311015358436
    ttmedian 7.267012239287151
    slnshift 3.6335061196435756
    slnmean 1.2901980553014483
    slnsd 0.2
  ]
  edge
  [
    source 67
    target 66
    length 151.97090261943515
    ttmedian 16.885655846603907
    slnshift 8.442827923301953
    slnmean 2.1333173144897297
    slnsd 0.2
  ]
  edge
  [
    source 68
    target 69
    length 133.76721811091065
    ttmedian 14.863024234545627
    slnshift 7.431512117272813
    slnmean 2.005729353134019
    slnsd 0.2
  ]
  edge
  [
    source 68
    target 78
    length 96.73998262843631
    ttmedian 10.748886958715145
    slnshift 5.374443479357573
    slnmean 1.6816550299290502
    slnsd 0.2
  ]
  edge
  [
    source 68
    target 58
    length 93.00085938737055
    ttmedian 10.33342882081895
    slnshift 5.166714410409475
    slnmean 1.6422369759388105
    slnsd 0.2
  ]
  edge
  [
    source 68
    target 67
    length 78.79007850059429
    ttmedian 8.754453166732699
    slnshift 4.377226583366349
    slnmean 1.4764153236878081
    slnsd 0.2
  ]
  edge
  [
    source 69
    target 79
    length 115.83657185893476
    ttmedian 12.870730206548307
    slnshift 6.435365103274154
    slnmean 1.861808576544142
    slnsd 0.2
  ]
  edge
  [
    source 69
    target 59
    length 101.73491457984885
    ttmedian 11.303879397760983
    slnshift 5.651939698880492
    slnmean 1.7319987957774852
    slnsd 0.2
  ]
  edge
  [
    source 69
    target 68
    length 142.33540252397347
    ttmedian 15.815044724885942
    slnshift 7.907522362442971
    slnmean 2.0678145041978375
    slnsd 0.2
  ]
  edge
  [
    source 70
    target 71
    length 84.23548607358572
    ttmedian 9.359498452620635
    slnshift 4.679749226310317
    slnmean 1.543244524372148
    slnsd 0.2
  ]
  edge
  [
    source 70
    target 80
    length 114.68120720637054
    ttmedian 12.742356356263393
    slnshift 6.3711781781316965
    slnmean 1.851784409802093
    slnsd 0.2
  ]
  edge
  [
    source 70
    target 60
    length 99.09456057026406
    ttmedian 11.010506730029341
    slnshift 5.505253365014671
    slnmean 1.7057027936413731
    slnsd 0.2
  ]
  edge
  [
    source 71
    target 72
    length 80.91636761730847
    ttmedian 8.990707513034273
    slnshift 4.495353756517137
    slnmean 1.5030443648307132
    slnsd 0.2
  ]
  edge
  [
    source 71
    target 81
    length 160.89406972108733
    ttmedian 17.877118857898594
    slnshift 8.938559428949297
    slnmean 2.1903744384999277
    slnsd 0.2
  ]
  edge
  [
    source 71
    target 61
    length 107.24023814441502
    ttmedian 11.915582016046113
    slnshift 5.957791008023056
    slnmean 1.7846997761369294
    slnsd 0.2
  ]
  edge
  [
    source 71
    target 70
    length 116.7732256148204
    ttmedian 12.974802846091157
    slnshift 6.487401423045578
    slnmean 1.869862053470294
    slnsd 0.2
  ]
  edge
  [
    source 72
    target 73
    length 76.55644094050275
    ttmedian 8.506271215611417
    slnshift 4.253135607805708
    slnmean 1.4476565009749558
    slnsd 0.2
  ]
  edge
  [
    source 72
    target 82
    length 83.47231354909394
    ttmedian 9.274701505454882
    slnshift 4.637350752727441
    slnmean 1.534143244715397
    slnsd 0.2
  ]
  edge
  [
    source 72
    target 62
    length 102.12576815199381
    ttmedian 11.347307572443757
    slnshift 5.6736537862218785
    slnmean 1.7358333169398645
    slnsd 0.2
  ]
  edge
  [
    source 72
    target 71
    length 86.40593770029837
    ttmedian 9.600659744477596
    slnshift 4.800329872238798
    slnmean 1.5686846389355846
    slnsd 0.2
  ]
  edge
  [
    source 73
    target 74
    length 116.24217770873958
    ttmedian 12.915797523193286
    slnshift 6.457898761596643
    slnmean 1.865303995780355
    slnsd 0.2
  ]
  edge
  [
    source 73
    target 83
    length 111.58219517545145
    ttmedian 12.398021686161272
    slnshift 6.199010843080636
    slnmean 1.8243897378843497
    slnsd 0.2
  ]
  edge
  [
    source 73
    target 63
    length 82.2352783837932
    ttmedian 9.1372531537548
    slnshift 4.5686265768774
    slnmean 1.5192126295140922
    slnsd 0.2
  ]
  edge
  [
    source 73
    target 72
    length 104.94161858309576
    ttmedian 11.660179842566196
    slnshift 5.830089921283098
    slnmean 1.7630324241342084
    slnsd 0.2
  ]
  edge
  [
    source 74
    target 75
    length 100.10067344142183
    ttmedian 11.122297049046871
    slnshift 5.5611485245234356
    slnmean 1.7158046560889113
    slnsd 0.2
  ]
  edge
  [
    source 74
    target 84
    length 95.98110635714828
    ttmedian 10.664567373016475
    slnshift 5.332283686508237
    slnmean 1.6737796054225766
    slnsd 0.2
  ]
  edge
  [
    source 74
    target 64
    length 75.0359340818758
    ttmedian 8.33732600909731
    slnshift 4.168663004548655
    slnmean 1.4275953619899615
    slnsd 0.2
  ]
  edge
  [
    source 74
    target 73
    length 81.14949411531127
    ttmedian 9.016610457256808
    slnshift 4.508305228628404
    slnmean 1.505921302099364
    slnsd 0.2
  ]
  edge
  [
    source 75
    target 76
    length 105.40213247180826
    ttmedian 11.71134805242314
    slnshift 5.85567402621157
    slnmean 1.7674111101867735
    slnsd 0.2
  ]
  edge
  [
    source 75
    target 85
    length 128.85224285698462
    ttmedian 14.31691587299829
    slnshift 7.158457936499145
    slnmean 1.968294585779349
    slnsd 0.2
  ]
  edge
  [
    source 75
    target 65
    length 78.46249424088258
    ttmedian 8.71805491565362
    slnshift 4.35902745782681
    slnmean 1.472248972347866
    slnsd 0.2
  ]
  edge
  [
    source 75
    target 74
    length 135.56694478589924
    ttmedian 15.062993865099916
    slnshift 7.531496932549958
    slnmean 2.0190938178778985
    slnsd 0.2
  ]
  edge
  [
    source 76
    target 77
    length 104.67203969941548
    ttmedian 11.630226633268386
    slnshift 5.815113316634193
    slnmean 1.7604602727341891
    slnsd 0.2
  ]
  edge
  [
    source 76
    target 86
    length 72.79995143038457
    ttmedian 8.088883492264952
    slnshift 4.044441746132476
    slnmean 1.3973435301412058
    slnsd 0.2
  ]
  edge
  [
    source 76
    target 66
    length 108.42055235645533
    ttmedian 12.046728039606148
    slnshift 6.023364019803074
    slnmean 1.7956459105223428
    slnsd 0.2
  ]
  edge
  [
    source 76
    target 75
    length 124.78770087394125
    ttmedian 13.865300097104583
    slnshift 6.932650048552292
    slnmean 1.936242142493151
    slnsd 0.2
  ]
  edge
  [
    source 77
    target 78
    length 106.8332127402717
    ttmedian 11.8703569711413
    slnshift 5.93517848557065
    slnmean 1.7808971009996044
    slnsd 0.2
  ]
  edge
  [
    source 77
    target 87
    length 178.56488121894856
    ttmedian 19.84054235766095
    slnshift 9.920271178830475
    slnmean 2.294580257498797
    slnsd 0.2
  ]
  edge
  [
    source 77
    target 67
    length 108.02223702233753
    ttmedian 12.002470780259726
    slnshift 6.001235390129863
    slnmean 1.791965346388875
    slnsd 0.2
  ]
  edge
  [
    source 77
    target 76
    length 124.99968207098733
    ttmedian 13.888853563443037
    slnshift 6.944426781721519
    slnmean 1.9379394359708007
    slnsd 0.2
  ]
  edge
  [
    source 78
    target 79
    length 100.27594745541309
    ttmedian 11.141771939490344
    slnshift 5.570885969745172
    slnmean 1.7175541022858762
    slnsd 0.2
  ]
  edge
  [
    source 78
    target 88
    length 127.76449903030876
    ttmedian 14.196055447812084
    slnshift 7.098027723906042
    slnmean 1.9598169600924225
    slnsd 0.2
  ]
  edge
  [
    source 78
    target 68
    length 54.94195051302071
    ttmedian 6.104661168113412
    slnshift 3.052330584056706
    slnmean 1.115905424743977
    slnsd 0.2
  ]
  edge
  [
    source 78
    target 77
    length 80.70411434842576
    ttmedian 8.967123816491751
    slnshift 4.483561908245876
    slnmean 1.5004177993320047
    slnsd 0.2
  ]
  edge
  [
    source 79
    target 89
    length 91.29542386878565
    ttmedian 10.143935985420628
    slnshift 5.071967992710314
    slnmean 1.623728906529507
    slnsd 0.2
  ]
  edge
  [
    source 79
    target 69
    length 113.42614398090238
    ttmedian 12.60290488676693
    slnshift 6.301452443383465
    slnmean 1.8407801533945147
    slnsd 0.2
  ]
  edge
  [
    source 79
    target 78
    length 104.82088009961434
    ttmedian 11.646764455512704
    slnshift 5.823382227756352
    slnmean 1.7618812317373438
    slnsd 0.2
  ]
  edge
  [
    source 80
    target 81
    length 131.16773898645783
    ttmedian 14.574193220717538
    slnshift 7.287096610358769
    slnmean 1.9861051965059506
    slnsd 0.2
  ]
  edge
  [
    source 80
    target 90
    length 109.33373054462072
    ttmedian 12.148192282735636
    slnshift 6.074096141367818
    slnmean 1.8040331948433375
    slnsd 0.2
  ]
  edge
  [
    source 80
    target 70
    length 105.31992029464546
    ttmedian 11.702213366071717
    slnshift 5.851106683035859
    slnmean 1.7666308199394571
    slnsd 0.2
  ]
  edge
  [
    source 81
    target 82
    length 96.43836037421927
    ttmedian 10.715373374913252
    slnshift 5.357686687456626
    slnmean 1.6785322937606573
    slnsd 0.2
  ]
  edge
  [
    source 81
    target 91
    length 104.18271672068092
    ttmedian 11.57585741340899
    slnshift 5.787928706704495
    slnmean 1.7557744912615345
    slnsd 0.2
  ]
  edge
  [
    source 81
    target 71
    length 123.04477420827378
    ttmedian 13.671641578697086
    slnshift 6.835820789348543
    slnmean 1.922176549191237
    slnsd 0.2
  ]
  edge
  [
    source 81
    target 80
    length 120.7829551116614
    ttmedian 13.420328345740154
    slnshift 6.710164172870077
    slnmean 1.9036234175799704
    slnsd 0.2
  ]
  edge
  [
    source 82
    target 83
    length 124.59165236299907
    ttmedian 13.843516929222119
    slnshift 6.9217584646110595
    slnmean 1.9346698507313422
    slnsd 0.2
  ]
  edge
  [
    source 82
    target 92
    length 94.17748226937077
    ttmedian 10.464164696596752
    slnshift 5.232082348298376
    slnmean 1.6548093533847004
    slnsd 0.2
  ]
  edge
  [
    source 82
    target 72
    length 84.04950955132718
    ttmedian 9.338834394591908
    slnshift 4.669417197295954
    slnmean 1.5410342667400927
    slnsd 0.2
  ]
  edge
  [
    source 82
    target 81
    length 88.88293485856997
    ttmedian 9.875881650952218
    slnshift 4.937940825476109
    slnmean 1.5969484073510045
    slnsd 0.2
  ]
  edge
  [
    source 83
    target 84
    length 109.12994063656289
    ttmedian 12.125548959618099
    slnshift 6.0627744798090495
    slnmean 1.8021675302587683
    slnsd 0.2
  ]
  edge
  [
    source 83
    target 93
    length 149.6103891750578
    ttmedian 16.62337657500642
    slnshift 8.31168828750321
    slnmean 2.1176627515902635
    slnsd 0.2
  ]
  edge
  [
    source 83
    target 73
    length 106.1661163268251
    ttmedian 11.796235147425012
    slnshift 5.898117573712506
    slnmean 1.7746332446997366
    slnsd 0.2
  ]
  edge
  [
    source 83
    target 82
    length 112.89078456016807
    ttmedian 12.543420506685342
    slnshift 6.271710253342671
    slnmean 1.8360490851287465
    slnsd 0.2
  ]
  edge
  [
    source 84
    target 85
    length 66.05706629972732
    ttmedian 7.339674033303036
    slnshift 3.669837016651518
    slnmean 1.300147251448603
    slnsd 0.2
  ]
  edge
  [
    source 84
    target 94
    length 195.41935012536763
    ttmedian 21.713261125040848
    slnshift 10.856630562520424
    slnmean 2.3847760050867555
    slnsd 0.2
  ]
  edge
  [
    source 84
    target 74
    length 75.89035894210481
    ttmedian 8.432262104678312
    slnshift 4.216131052339156
    slnmean 1.4389178952766202
    slnsd 0.2
  ]
  edge
  [
    source 84
    target 83
    length 157.52998768884174
    ttmedian 17.50333196542686
    slnshift 8.75166598271343
    slnmean 2.1692440802705115
    slnsd 0.2
  ]
  edge
  [
    source 85
    target 86
    length 87.0217729013407
    ttmedian 9.669085877926744
    slnshift 4.834542938963372
    slnmean 1.5757865926816008
    slnsd 0.2
  ]
  edge
  [
    source 85
    target 95
    length 127.05219622355692
    ttmedian 14.116910691506325
    slnshift 7.0584553457531625
    slnmean 1.9542262380138482
    slnsd 0.2
  ]
  edge
  [
    source 85
    target 75
    length 122.60592574976523
    ttmedian 13.622880638862803
    slnshift 6.8114403194314015
    slnmean 1.9186035984490313
    slnsd 0.2
  ]
  edge
  [
    source 85
    target 84
    length 101.62098264586466
    ttmedian 11.291220293984962
    slnshift 5.645610146992481
    slnmean 1.7308782780304297
    slnsd 0.2
  ]
  edge
  [
    source 86
    target 87
    length 95.92521170446794
    ttmedian 10.658356856051993
    slnshift 5.329178428025997
    slnmean 1.6731970852134928
    slnsd 0.2
  ]
  edge
  [
    source 86
    target 96
    length 79.04305802192853
    ttmedian 8.782562002436503
    slnshift 4.3912810012182515
    slnmean 1.4796209843437174
    slnsd 0.2
  ]
  edge
  [
    source 86
    target 76
    length 98.66907365962719
    ttmedian 10.963230406625243
    slnshift 5.4816152033126215
    slnmean 1.701399802661152
    slnsd 0.2
  ]
  edge
  [
    source 86
    target 85
    length 131.69046895577864
    ttmedian 14.632274328419848
    slnshift 7.316137164209924
    slnmean 1.9900824788780078
    slnsd 0.2
  ]
  edge
  [
    source 87
    target 88
    length 109.74953580064815
    ttmedian 12.194392866738683
    slnshift 6.097196433369342
    slnmean 1.8078290644433206
    slnsd 0.2
  ]
  edge
  [
    source 87
    target 97
    length 80.18247214789402
    ttmedian 8.909163571988223
    slnshift 4.454581785994112
    slnmean 1.4939331813184902
    slnsd 0.2
  ]
  edge
  [
    source 87
    target 77
    length 133.57766632635582
    ttmedian 14.841962925150646
    slnshift 7.420981462575323
    slnmean 2.0043113210163175
    slnsd 0.2
  ]
  edge
  [
    source 87
    target 86
    length 74.01418863159441
    ttmedian 8.223798736843824
    slnshift 4.111899368421912
    slnmean 1.4138850551933626
    slnsd 0.2
  ]
  edge
  [
    source 88
    target 89
    length 105.38601807323408
    ttmedian 11.709557563692677
    slnshift 5.854778781846338
    slnmean 1.7672582135598627
    slnsd 0.2
  ]
  edge
  [
    source 88
    target 98
    length 80.02775047450565
    ttmedian 8.891972274945072
    slnshift 4.445986137472536
    slnmean 1.4920016975597565
    slnsd 0.2
  ]
  edge
  [
    source 88
    target 78
    length 82.22819509247915
    ttmedian 9.136466121386572
    slnshift 4.568233060693286
    slnmean 1.5191264913440936
    slnsd 0.2
  ]
  edge
  [
    source 88
    target 87
    length 130.98762530741251
    ttmedian 14.554180589712502
    slnshift 7.277090294856251
    slnmean 1.9847310975408536
    slnsd 0.2
  ]
  edge
  [
    source 89
    target 99
    length 89.07333013527881
    ttmedian 9.897036681697646
    slnshift 4.948518340848823
    slnmean 1.5990882066928596
    slnsd 0.2
  ]
  edge
  [
    source 89
    target 79
    length 68.47038158557923
    ttmedian 7.60782017617547
    slnshift 3.803910088087735
    slnmean 1.3360295082544196
    slnsd 0.2
  ]
  edge
  [
    source 89
    target 88
    length 87.39958663238794
    ttmedian 9.711065181376437
    slnshift 4.8555325906882185
    slnmean 1.580118795147595
    slnsd 0.2
  ]
  edge
  [
    source 90
    target 91
    length 120.82287744829084
    ttmedian 13.424764160921205
    slnshift 6.7123820804606025
    slnmean 1.9039538925239714
    slnsd 0.2
  ]
  edge
  [
    source 90
    target 80
    length 83.33097929792801
    ttmedian 9.258997699769779
    slnshift 4.6294988498848895
    slnmean 1.5324486224741138
    slnsd 0.2
  ]
  edge
  [
    source 91
    target 92
    length 135.48566596775032
    ttmedian 15.053962885305591
    slnshift 7.5269814426527955
    slnmean 2.0184940906119695
    slnsd 0.2
  ]
  edge
  [
    source 91
    target 81
    length 128.66870564646337
    ttmedian 14.29652284960704
    slnshift 7.14826142480352
    slnmean 1.9668691697682328
    slnsd 0.2
  ]
  edge
  [
    source 91
    target 90
    length 101.84896068176059
    ttmedian 11.316551186862288
    slnshift 5.658275593431144
    slnmean 1.7331191803236088
    slnsd 0.2
  ]
  edge
  [
    source 92
    target 93
    length 134.10234664396293
    ttmedian 14.900260738218103
    slnshift 7.4501303691090515
    slnmean 2.0082315314476276
    slnsd 0.2
  ]
  edge
  [
    source 92
    target 82
    length 159.31893965004716
    ttmedian 17.702104405560796
    slnshift 8.851052202780398
    slnmean 2.1805363449221655
    slnsd 0.2
  ]
  edge
  [
    source 92
    target 91
    length 60.90169833345435
    ttmedian 6.766855370383817
    slnshift 3.3834276851919083
    slnmean 1.2188893036796844
    slnsd 0.2
  ]
  edge
  [
    source 93
    target 94
    length 133.5321694554024
    ttmedian 14.836907717266932
    slnshift 7.418453858633466
    slnmean 2.0039706606189926
    slnsd 0.2
  ]
  edge
  [
    source 93
    target 83
    length 116.45334871093722
    ttmedian 12.939260967881914
    slnshift 6.469630483940957
    slnmean 1.8671189946625177
    slnsd 0.2
  ]
  edge
  [
    source 93
    target 92
    length 92.2377864426502
    ttmedian 10.248642938072244
    slnshift 5.124321469036122
    slnmean 1.6339981199736753
    slnsd 0.2
  ]
  edge
  [
    source 94
    target 95
    length 118.0616077411218
    ttmedian 13.1179564156802
    slnshift 6.5589782078401
    slnmean 1.8808348298240085
    slnsd 0.2
  ]
  edge
  [
    source 94
    target 84
    length 110.42667038713961
    ttmedian 12.269630043015512
    slnshift 6.134815021507756
    slnmean 1.8139799263648348
    slnsd 0.2
  ]
  edge
  [
    source 94
    target 93
    length 86.58624513288575
    ttmedian 9.620693903653972
    slnshift 4.810346951826986
    slnmean 1.5707692128732904
    slnsd 0.2
  ]
  edge
  [
    source 95
    target 96
    length 145.07121851431313
    ttmedian 16.119024279368126
    slnshift 8.059512139684063
    slnmean 2.0868530261114304
    slnsd 0.2
  ]
  edge
  [
    source 95
    target 85
    length 103.56836258041722
    ttmedian 11.507595842268579
    slnshift 5.753797921134289
    slnmean 1.7498601447931952
    slnsd 0.2
  ]
  edge
  [
    source 95
    target 94
    length 134.5921876757564
    ttmedian 14.95468751952849
    slnshift 7.477343759764245
    slnmean 2.011877616588088
    slnsd 0.2
  ]
  edge
  [
    source 96
    target 97
    length 94.04128040367178
    ttmedian 10.449031155963532
    slnshift 5.224515577981766
    slnmean 1.653362081204832
    slnsd 0.2
  ]
  edge
  [
    source 96
    target 86
    length 66.4382889754849
    ttmedian 7.382032108387211
    slnshift 3.6910160541936055
    slnmean 1.3059017735896365
    slnsd 0.2
  ]
  edge
  [
    source 96
    target 95
    length 106.31110072073562
    ttmedian 11.81234452452618
    slnshift 5.90617226226309
    slnmean 1.775997950228423
    slnsd 0.2
  ]
  edge
  [
    source 97
    target 98
    length 94.29230910157632
    ttmedian 10.47692323350848
    slnshift 5.23846161675424
    slnmean 1.6560278706396732
    slnsd 0.2
  ]
  edge
  [
    source 97
    target 87
    length 72.96378590568894
    ttmedian 8.107087322854326
    slnshift 4.053543661427163
    slnmean 1.3995914766783915
    slnsd 0.2
  ]
  edge
  [
    source 97
    target 96
    length 65.5083885426125
    ttmedian 7.278709838068055
    slnshift 3.6393549190340275
    slnmean 1.2918064458977245
    slnsd 0.2
  ]
  edge
  [
    source 98
    target 99
    length 116.66223934096934
    ttmedian 12.962471037885482
    slnshift 6.481235518942741
    slnmean 1.868911158693145
    slnsd 0.2
  ]
  edge
  [
    source 98
    target 88
    length 137.3583228195292
    ttmedian 15.262035868836579
    slnshift 7.631017934418289
    slnmean 2.03222124851025
    slnsd 0.2
  ]
  edge
  [
    source 98
    target 97
    length 141.81234650256843
    ttmedian 15.75692738917427
    slnshift 7.878463694587135
    slnmean 2.0641329222462605
    slnsd 0.2
  ]
  edge
  [
    source 99
    target 89
    length 114.28855408582525
    ttmedian 12.698728231758361
    slnshift 6.3493641158791805
    slnmean 1.8483546686587076
    slnsd 0.2
  ]
  edge
  [
    source 99
    target 98
    length 100.10260356544907
    ttmedian 11.122511507272119
    slnshift 5.561255753636059
    slnmean 1.7158239377316111
    slnsd 0.2
  ]
]
