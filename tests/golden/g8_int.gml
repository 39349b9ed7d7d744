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
    name "(-71.099, 42.3)"
    lng -71.099
    lat 42.3
  ]
  node
  [
    id 9
    name "(-71.099, 42.301)"
    lng -71.099
    lat 42.301
  ]
  node
  [
    id 10
    name "(-71.099, 42.302)"
    lng -71.099
    lat 42.302
  ]
  node
  [
    id 11
    name "(-71.099, 42.303)"
    lng -71.099
    lat 42.303
  ]
  node
  [
    id 12
    name "(-71.099, 42.304)"
    lng -71.099
    lat 42.304
  ]
  node
  [
    id 13
    name "(-71.099, 42.305)"
    lng -71.099
    lat 42.305
  ]
  node
  [
    id 14
    name "(-71.099, 42.306)"
    lng -71.099
    lat 42.306
  ]
  node
  [
    id 15
    name "(-71.099, 42.307)"
    lng -71.099
    lat 42.307
  ]
  node
  [
    id 16
    name "(-71.098, 42.3)"
    lng -71.098
    lat 42.3
  ]
  node
  [
    id 17
    name "(-71.098, 42.301)"
    lng -71.098
    lat 42.301
  ]
  node
  [
    id 18
    name "(-71.098, 42.302)"
    lng -71.098
    lat 42.302
  ]
  node
  [
    id 19
    name "(-71.098, 42.303)"
    lng -71.098
    lat 42.303
  ]
  node
  [
    id 20
    name "(-71.098, 42.304)"
    lng -71.098
    lat 42.304
  ]
  node
  [
    id 21
    name "(-71.098, 42.305)"
    lng -71.098
    lat 42.305
  ]
  node
  [
    id 22
    name "(-71.098, 42.306)"
    lng -71.098
    lat 42.306
  ]
  node
  [
    id 23
    name "(-71.098, 42.307)"
    lng -71.098
    lat 42.307
  ]
  node
  [
    id 24
    name "(-71.097, 42.3)"
    lng -71.097
    lat 42.3
  ]
  node
  [
    id 25
    name "(-71.097, 42.301)"
    lng -71.097
    lat 42.301
  ]
  node
  [
    id 26
    name "(-71.097, 42.302)"
    lng -71.097
    lat 42.302
  ]
  node
  [
    id 27
    name "(-71.097, 42.303)"
    lng -71.097
    lat 42.303
  ]
  node
  [
    id 28
    name "(-71.097, 42.304)"
    lng -71.097
    lat 42.304
  ]
  node
  [
    id 29
    name "(-71.097, 42.305)"
    lng -71.097
    lat 42.305
  ]
  node
  [
    id 30
    name "(-71.097, 42.306)"
    lng -71.097
    lat 42.306
  ]
  node
  [
    id 31
    name "(-71.097, 42.307)"
    lng -71.097
    lat 42.307
  ]
  node
  [
    id 32
    name "(-71.096, 42.3)"
    lng -71.096
    lat 42.3
  ]
  node
  [
    id 33
    name "(-71.096, 42.301)"
    lng -71.096
    lat 42.301
  ]
  node
  [
    id 34
    name "(-71.096, 42.302)"
    lng -71.096
    lat 42.302
  ]
  node
  [
    id 35
    name "(-71.096, 42.303)"
    lng -71.096
    lat 42.303
  ]
  node
  [
    id 36
    name "(-71.096, 42.304)"
    lng -71.096
    lat 42.304
  ]
  node
  [
    id 37
    name "(-71.096, 42.305)"
    lng -71.096
    lat 42.305
  ]
  node
  [
    id 38
    name "(-71.096, 42.306)"
    lng -71.096
    lat 42.306
  ]
  node
  [
    id 39
    name "(-71.096, 42.307)"
    lng -71.096
    lat 42.307
  ]
  node
  [
    id 40
    name "(-71.095, 42.3)"
    lng -71.095
    lat 42.3
  ]
  node
  [
    id 41
    name "(-71.095, 42.301)"
    lng -71.095
    lat 42.301
  ]
  node
  [
    id 42
    name "(-71.095, 42.302)"
    lng -71.095
    lat 42.302
  ]
  node
  [
    id 43
    name "(-71.095, 42.303)"
    lng -71.095
    lat 42.303
  ]
  node
  [
    id 44
    name "(-71.095, 42.304)"
    lng -71.095
    lat 42.304
  ]
  node
  [
    id 45
    name "(-71.095, 42.305)"
    lng -71.095
    lat 42.305
  ]
  node
  [
    id 46
    name "(-71.095, 42.306)"
    lng -71.095
    lat 42.306
  ]
  node
  [
    id 47
    name "(-71.095, 42.307)"
    lng -71.095
    lat 42.307
  ]
  node
  [
    id 48
    name "(-71.094, 42.3)"
    lng -71.094
    lat 42.3
  ]
  node
  [
    id 49
    name "(-71.094, 42.301)"
    lng -71.094
    lat 42.301
  ]
  node
  [
    id 50
    name "(-71.094, 42.302)"
    lng -71.094
    lat 42.302
  ]
  node
  [
    id 51
    name "(-71.094, 42.303)"
    lng -71.094
    lat 42.303
  ]
  node
  [
    id 52
    name "(-71.094, 42.304)"
    lng -71.094
    lat 42.304
  ]
  node
  [
    id 53
    name "(-71.094, 42.305)"
    lng -71.094
    lat 42.305
  ]
  node
  [
    id 54
    name "(-71.094, 42.306)"
    lng -71.094
    lat 42.306
  ]
  node
  [
    id 55
    name "(-71.094, 42.307)"
    lng -71.094
    lat 42.307
  ]
  node
  [
    id 56
    name "(-71.093, 42.3)"
    lng -71.093
    lat 42.3
  ]
  node
  [
    id 57
    name "(-71.093, 42.301)"
    lng -71.093
    lat 42.301
  ]
  node
  [
    id 58
    name "(-71.093, 42.302)"
    lng -71.093
    lat 42.302
  ]
  node
  [
    id 59
    name "(-71.093, 42.303)"
    lng -71.093
    lat 42.303
  ]
  node
  [
    id 60
    name "(-71.093, 42.304)"
    lng -71.093
    lat 42.304
  ]
  node
  [
    id 61
    name "(-71.093, 42.305)"
    lng -71.093
    lat 42.305
  ]
  node
  [
    id 62
    name "(-71.093, 42.306)"
    lng -71.093
    lat 42.306
  ]
  node
  [
    id 63
    name "(-71.093, 42.307)"
    lng -71.093
    lat 42.307
  ]
  edge
  [
    source 0
    target 1
    length 230.0
    ttmedian 23.0
    slnshift 11.5
    slnmean 2.4423470353692043
    slnsd 0.2
  ]
  edge
  [
    source 0
    target 8
    length 400.0
    ttmedian 40.0
    slnshift 20.0
    slnmean 2.995732273553991
    slnsd 0.2
  ]
  edge
  [
    source 1
    target 2
    length 250.0
    ttmedian 25.0
    slnshift 12.5
    slnmean 2.5257286443082556
    slnsd 0.2
  ]
  edge
  [
    source 1
    target 9
    length 460.0
    ttmedian 46.0
    slnshift 23.0
    slnmean 3.1354942159291497
    slnsd 0.2
  ]
  edge
  [
    source 2
    target 3
    length 280.0
    ttmedian 28.0
    slnshift 14.0
    slnmean 2.6390573296152584
    slnsd 0.2
  ]
  edge
  [
    source 2
    target 10
    length 390.0
    ttmedian 39.0
    slnshift 19.5
    slnmean 2.970414465569701
    slnsd 0.2
  ]
  edge
  [
    source 3
    target 4
    length 600.0
    ttmedian 60.0
    slnshift 30.0
    slnmean 3.4011973816621555
    slnsd 0.2
  ]
  edge
  [
    source 3
    target 11
    length 410.0
    ttmedian 41.0
    slnshift 20.5
    slnmean 3.0204248861443626
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
    target 12
    length 470.0
    ttmedian 47.0
    slnshift 23.5
    slnmean 3.1570004211501135
    slnsd 0.2
  ]
  edge
  [
    source 5
    target 6
    length 330.0
    ttmedian 33.0
    slnshift 16.5
    slnmean 2.803360380906535
    slnsd 0.2
  ]
  edge
  [
    source 5
    target 13
    length 290.0
    ttmedian 29.0
    slnshift 14.5
    slnmean 2.6741486494265287
    slnsd 0.2
  ]
  edge
  [
    source 6
    target 7
    length 350.0
    ttmedian 35.0
    slnshift 17.5
    slnmean 2.8622008809294686
    slnsd 0.2
  ]
  edge
  [
    source 6
    target 14
    length 490.0
    ttmedian 49.0
    slnshift 24.5
    slnmean 3.1986731175506815
    slnsd 0.2
  ]
  edge
  [
    source 7
    target 15
    length 540.0
    ttmedian 54.0
    slnshift 27.0
    slnmean 3.295836866004329
    slnsd 0.2
  ]
  edge
  [
    source 8
    target 9
    length 380.0
    ttmedian 38.0
    slnshift 19.0
    slnmean 2.9444389791664403
    slnsd 0.2
  ]
  edge
  [
    source 8
    target 16
    length 340.0
    ttmedian 34.0
    slnshift 17.0
    slnmean 2.833213344056216
    slnsd 0.2
  ]
  edge
  [
    source 9
    target 10
    length 580.0
    ttmedian 58.0
    slnshift 29.0
    slnmean 3.367295829986474
    slnsd 0.2
  ]
  edge
  [
    source 9
    target 17
    length 490.0
    ttmedian 49.0
    slnshift 24.5
    slnmean 3.1986731175506815
    slnsd 0.2
  ]
  edge
  [
    source 10
    target 11
    length 290.0
    ttmedian 29.0
    slnshift 14.5
    slnmean 2.6741486494265287
    slnsd 0.2
  ]
  edge
  [
    source 10
    target 18
    length 240.0
    ttmedian 24.0
    slnshift 12.0
    slnmean 2.4849066497880004
    slnsd 0.2
  ]
  edge
  [
    source 11
    target 12
    length 270.0
    ttmedian 27.0
    slnshift 13.5
    slnmean 2.6026896854443837
    slnsd 0.2
  ]
  edge
  [
    source 11
    target 19
    length 460.0
    ttmedian 46.0
    slnshift 23.0
    slnmean 3.1354942159291497
    slnsd 0.2
  ]
  edge
  [
    source 12
    target 13
    length 220.0
    ttmedian 22.0
    slnshift 11.0
    slnmean 2.3978952727983707
    slnsd 0.2
  ]
  edge
  [
    source 12
    target 20
    length 580.0
    ttmedian 58.0
    slnshift 29.0
    slnmean 3.367295829986474
    slnsd 0.2
  ]
  edge
  [
    source 13
    target 14
    length 290.0
    ttmedian 29.0
    slnshift 14.5
    slnmean 2.6741486494265287
    slnsd 0.2
  ]
  edge
  [
    source 13
    target 21
    length 570.0
    ttmedian 57.0
    slnshift 28.5
    slnmean 3.349904087274605
    slnsd 0.2
  ]
  edge
  [
    source 14
    target 15
    length 370.0
    ttmedian 37.0
    slnshift 18.5
    slnmean 2.917770732084279
    slnsd 0.2
  ]
  edge
  [
    source 14
    target 22
    length 350.0
    ttmedian 35.0
    slnshift 17.5
    slnmean 2.8622008809294686
    slnsd 0.2
  ]
  edge
  [
    source 15
    target 23
    length 480.0
    ttmedian 48.0
    slnshift 24.0
    slnmean 3.1780538303479458
    slnsd 0.2
  ]
  edge
  [
    source 16
    target 17
    length 470.0
    ttmedian 47.0
    slnshift 23.5
    slnmean 3.1570004211501135
    slnsd 0.2
  ]
  edge
  [
    source 16
    target 24
    length 420.0
    ttmedian 42.0
    slnshift 21.0
    slnmean 3.044522437723423
    slnsd 0.2
  ]
  edge
  [
    source 17
    target 18
    length 510.0
    ttmedian 51.0
    slnshift 25.5
    slnmean 3.2386784521643803
    slnsd 0.2
  ]
  edge
  [
    source 17
    target 25
    length 400.0
    ttmedian 40.0
    slnshift 20.0
    slnmean 2.995732273553991
    slnsd 0.2
  ]
  edge
  [
    source 18
    target 19
    length 410.0
    ttmedian 41.0
    slnshift 20.5
    slnmean 3.0204248861443626
    slnsd 0.2
  ]
  edge
  [
    source 18
    target 26
    length 280.0
    ttmedian 28.0
    slnshift 14.0
    slnmean 2.6390573296152584
    slnsd 0.2
  ]
  edge
  [
    source 19
    target 20
    length 480.0
    ttmedian 48.0
    slnshift 24.0
    slnmean 3.1780538303479458
    slnsd 0.2
  ]
  edge
  [
    source 19
    target 27
    length 480.0
    ttmedian 48.0
    slnshift 24.0
    slnmean 3.1780538303479458
    slnsd 0.2
  ]
  edge
  [
    source 20
    target 21
    length 510.0
    ttmedian 51.0
    slnshift 25.5
    slnmean 3.2386784521643803
    slnsd 0.2
  ]
  edge
  [
    source 20
    target 28
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
    length 200.0
    ttmedian 20.0
    slnshift 10.0
    slnmean 2.302585092994046
    slnsd 0.2
  ]
  edge
  [
    source 21
    target 29
    length 270.0
    ttmedian 27.0
    slnshift 13.5
    slnmean 2.6026896854443837
    slnsd 0.2
  ]
  edge
  [
    source 22
    target 23
    length 220.0
    ttmedian 22.0
    slnshift 11.0
    slnmean 2.3978952727983707
    slnsd 0.2
  ]
  edge
  [
    source 22
    target 30
    length 580.0
    ttmedian 58.0
    slnshift 29.0
    slnmean 3.367295829986474
    slnsd 0.2
  ]
  edge
  [
    source 23
    target 31
    length 270.0
    ttmedian 27.0
    slnshift 13.5
    slnmean 2.6026896854443837
    slnsd 0.2
  ]
  edge
  [
    source 24
    target 25
    length 510.0
    ttmedian 51.0
    slnshift 25.5
    slnmean 3.2386784521643803
    slnsd 0.2
  ]
  edge
  [
    source 24
    target 32
    length 560.0
    ttmedian 56.0
    slnshift 28.0
    slnmean 3.332204510175204
    slnsd 0.2
  ]
  edge
  [
    source 25
    target 26
    length 450.0
    ttmedian 45.0
    slnshift 22.5
    slnmean 3.1135153092103742
    slnsd 0.2
  ]
  edge
  [
    source 25
    target 33
    length 580.0
    ttmedian 58.0
    slnshift 29.0
    slnmean 3.367295829986474
    slnsd 0.2
  ]
  edge
  [
    source 26
    target 27
    length 470.0
    ttmedian 47.0
    slnshift 23.5
    slnmean 3.1570004211501135
    slnsd 0.2
  ]
  edge
  [
    source 26
    target 34
    length 240.0
    ttmedian 24.0
    slnshift 12.0
    slnmean 2.4849066497880004
    slnsd 0.2
  ]
  edge
  [
    source 27
    target 28
    length 570.0
    ttmedian 57.0
    slnshift 28.5
    slnmean 3.349904087274605
    slnsd 0.2
  ]
  edge
  [
    source 27
    target 35
    length 220.0
    ttmedian 22.0
    slnshift 11.0
    slnmean 2.3978952727983707
    slnsd 0.2
  ]
  edge
  [
    source 28
    target 29
    length 380.0
    ttmedian 38.0
    slnshift 19.0
    slnmean 2.9444389791664403
    slnsd 0.2
  ]
  edge
  [
    source 28
    target 36
    length 580.0
    ttmedian 58.0
    slnshift 29.0
    slnmean 3.367295829986474
    slnsd 0.2
  ]
  edge
  [
    source 29
    target 30
    length 420.0
    ttmedian 42.0
    slnshift 21.0
    slnmean 3.044522437723423
    slnsd 0.2
  ]
  edge
  [
    source 29
    target 37
    length 590.0
    ttmedian 59.0
    slnshift 29.5
    slnmean 3.3843902633457743
    slnsd 0.2
  ]
  edge
  [
    source 30
    target 31
    length 260.0
    ttmedian 26.0
    slnshift 13.0
    slnmean 2.5649493574615367
    slnsd 0.2
  ]
  edge
  [
    source 30
    target 38
    length 510.0
    ttmedian 51.0
    slnshift 25.5
    slnmean 3.2386784521643803
    slnsd 0.2
  ]
  edge
  [
    source 31
    target 39
    length 520.0
    ttmedian 52.0
    slnshift 26.0
    slnmean 3.258096538021482
    slnsd 0.2
  ]
  edge
  [
    source 32
    target 33
    length 290.0
    ttmedian 29.0
    slnshift 14.5
    slnmean 2.6741486494265287
    slnsd 0.2
  ]
  edge
  [
    source 32
    target 40
    length 540.0
    ttmedian 54.0
    slnshift 27.0
    slnmean 3.295836866004329
    slnsd 0.2
  ]
  edge
  [
    source 33
    target 34
    length 270.0
    ttmedian 27.0
    slnshift 13.5
    slnmean 2.6026896854443837
    slnsd 0.2
  ]
  edge
  [
    source 33
    target 41
    length 290.0
    ttmedian 29.0
    slnshift 14.5
    slnmean 2.6741486494265287
    slnsd 0.2
  ]
  edge
  [
    source 34
    target 35
    length 540.0
    ttmedian 54.0
    slnshift 27.0
    slnmean 3.295836866004329
    slnsd 0.2
  ]
  edge
  [
    source 34
    target 42
    length 580.0
    ttmedian 58.0
    slnshift 29.0
    slnmean 3.367295829986474
    slnsd 0.2
  ]
  edge
  [
    source 35
    target 36
    length 390.0
    ttmedian 39.0
    slnshift 19.5
    slnmean 2.970414465569701
    slnsd 0.2
  ]
  edge
  [
    source 35
    target 43
    length 290.0
    ttmedian 29.0
    slnshift 14.5
    slnmean 2.6741486494265287
    slnsd 0.2
  ]
  edge
  [
    source 36
    target 37
    length 360.0
    ttmedian 36.0
    slnshift 18.0
    slnmean 2.8903717578961645
    slnsd 0.2
  ]
  edge
  [
    source 36
    target 44
    length 340.0
    ttmedian 34.0
    slnshift 17.0
    slnmean 2.833213344056216
    slnsd 0.2
  ]
  edge
  [
    source 37
    target 38
    length 400.0
    ttmedian 40.0
    slnshift 20.0
    slnmean 2.995732273553991
    slnsd 0.2
  ]
  edge
  [
    source 37
    target 45
    length 540.0
    ttmedian 54.0
    slnshift 27.0
    slnmean 3.295836866004329
    slnsd 0.2
  ]
  edge
  [
    source 38
    target 39
    length 430.0
    ttmedian 43.0
    slnshift 21.5
    slnmean 3.068052935133617
    slnsd 0.2
  ]
  edge
  [
    source 38
    target 46
    length 540.0
    ttmedian 54.0
    slnshift 27.0
    slnmean 3.295836866004329
    slnsd 0.2
  ]
  edge
  [
    source 39
    target 47
    length 440.0
    ttmedian 44.0
    slnshift 22.0
    slnmean 3.091042453358316
    slnsd 0.2
  ]
  edge
  [
    source 40
    target 41
    length 260.0
    ttmedian 26.0
    slnshift 13.0
    slnmean 2.5649493574615367
    slnsd 0.2
  ]
  edge
  [
    source 40
    target 48
    length 480.0
    ttmedian 48.0
    slnshift 24.0
    slnmean 3.1780538303479458
    slnsd 0.2
  ]
  edge
  [
    source 41
    target 42
    length 250.0
    ttmedian 25.0
    slnshift 12.5
    slnmean 2.5257286443082556
    slnsd 0.2
  ]
  edge
  [
    source 41
    target 49
    length 220.0
    ttmedian 22.0
    slnshift 11.0
    slnmean 2.3978952727983707
    slnsd 0.2
  ]
  edge
  [
    source 42
    target 43
    length 390.0
    ttmedian 39.0
    slnshift 19.5
    slnmean 2.970414465569701
    slnsd 0.2
  ]
  edge
  [
    source 42
    target 50
    length 550.0
    ttmedian 55.0
    slnshift 27.5
    slnmean 3.3141860046725258
    slnsd 0.2
  ]
  edge
  [
    source 43
    target 44
    length 310.0
    ttmedian 31.0
    slnshift 15.5
    slnmean 2.740840023925201
    slnsd 0.2
  ]
  edge
  [
    source 43
    target 51
    length 310.0
    ttmedian 31.0
    slnshift 15.5
    slnmean 2.740840023925201
    slnsd 0.2
  ]
  edge
  [
    source 44
    target 45
    length 480.0
    ttmedian 48.0
    slnshift 24.0
    slnmean 3.1780538303479458
    slnsd 0.2
  ]
  edge
  [
    source 44
    target 52
    length 280.0
    ttmedian 28.0
    slnshift 14.0
    slnmean 2.6390573296152584
    slnsd 0.2
  ]
  edge
  [
    source 45
    target 46
    length 450.0
    ttmedian 45.0
    slnshift 22.5
    slnmean 3.1135153092103742
    slnsd 0.2
  ]
  edge
  [
    source 45
    target 53
    length 340.0
    ttmedian 34.0
    slnshift 17.0
    slnmean 2.833213344056216
    slnsd 0.2
  ]
  edge
  [
    source 46
    target 47
    length 590.0
    ttmedian 59.0
    slnshift 29.5
    slnmean 3.3843902633457743
    slnsd 0.2
  ]
  edge
  [
    source 46
    target 54
    length 350.0
    ttmedian 35.0
    slnshift 17.5
    slnmean 2.8622008809294686
    slnsd 0.2
  ]
  edge
  [
    source 47
    target 55
    length 390.0
    ttmedian 39.0
    slnshift 19.5
    slnmean 2.970414465569701
    slnsd 0.2
  ]
  edge
  [
    source 48
    target 49
    length 600.0
    ttmedian 60.0
    slnshift 30.0
    slnmean 3.4011973816621555
    slnsd 0.2
  ]
  edge
  [
    source 48
    target 56
    length 350.0
    ttmedian 35.0
    slnshift 17.5
    slnmean 2.8622008809294686
    slnsd 0.2
  ]
  edge
  [
    source 49
    target 50
    length 480.0
    ttmedian 48.0
    slnshift 24.0
    slnmean 3.1780538303479458
    slnsd 0.2
  ]
  edge
  [
    source 49
    target 57
    length 550.0
    ttmedian 55.0
    slnshift 27.5
    slnmean 3.3141860046725258
    slnsd 0.2
  ]
  edge
  [
    source 50
    target 51
    length 520.0
    ttmedian 52.0
    slnshift 26.0
    slnmean 3.258096538021482
    slnsd 0.2
  ]
  edge
  [
    source 50
    target 58
    length 360.0
    ttmedian 36.0
    slnshift 18.0
    slnmean 2.8903717578961645
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
    target 59
    length 530.0
    ttmedian 53.0
    slnshift 26.5
    slnmean 3.2771447329921766
    slnsd 0.2
  ]
  edge
  [
    source 52
    target 53
    length 220.0
    ttmedian 22.0
    slnshift 11.0
    slnmean 2.3978952727983707
    slnsd 0.2
  ]
  edge
  [
    source 52
    target 60
    length 260.0
    ttmedian 26.0
    slnshift 13.0
    slnmean 2.5649493574615367
    slnsd 0.2
  ]
  edge
  [
    source 53
    target 54
    length 270.0
    ttmedian 27.0
    slnshift 13.5
    slnmean 2.6026896854443837
    slnsd 0.2
  ]
  edge
  [
    source 53
    target 61
    length 390.0
    ttmedian 39.0
    slnshift 19.5
    slnmean 2.970414465569701
    slnsd 0.2
  ]
  edge
  [
    source 54
    target 55
    length 280.0
    ttmedian 28.0
    slnshift 14.0
    slnmean 2.6390573296152584
    slnsd 0.2
  ]
  edge
  [
    source 54
    target 62
    length 540.0
    ttmedian 54.0
    slnshift 27.0
    slnmean 3.295836866004329
    slnsd 0.2
  ]
  edge
  [
    source 55
    target 63
    length 420.0
    ttmedian 42.0
    slnshift 21.0
    slnmean 3.044522437723423
    slnsd 0.2
  ]
  edge
  [
    source 56
    target 57
    length 320.0
    ttmedian 32.0
    slnshift 16.0
    slnmean 2.772588722239781
    slnsd 0.2
  ]
  edge
  [
    source 57
    target 58
    length 480.0
    ttmedian 48.0
    slnshift 24.0
    slnmean 3.1780538303479458
    slnsd 0.2
  ]
  edge
  [
    source 58
    target 59
    length 410.0
    ttmedian 41.0
    slnshift 20.5
    slnmean 3.0204248861443626
    slnsd 0.2
  ]
  edge
  [
    source 59
    target 60
    length 550.0
    ttmedian 55.0
    slnshift 27.5
    slnmean 3.3141860046725258
    slnsd 0.2
  ]
  edge
  [
    source 60
    target 61
    length 490.0
    ttmedian 49.0
    slnshift 24.5
    slnmean 3.1986731175506815
    slnsd 0.2
  ]
  edge
  [
    source 61
    target 62
    length 490.0
    ttmedian 49.0
    slnshift 24.5
    slnmean 3.1986731175506815
    slnsd 0.2
  ]
  edge
  [
    source 62
    target 63
    length 200.0
    ttmedian 20.0
    slnshift 10.0
    slnmean 2.302585092994046
    slnsd 0.2
  ]
]
