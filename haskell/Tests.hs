-- The reference's test driver shape (Tests.hs): the four modules' HUnit values, exit 1 on any failure.
module Main (main) where

import qualified Data.BWT     as BWT
import qualified Data.FMIndex as FMIndex
import qualified Data.MTF     as MTF
import qualified Data.RLE     as RLE
import           System.Exit  (exitFailure, exitSuccess)
import           Test.HUnit

main :: IO ()
main = do
  c <- runTestTT (TestList [BWT.tests, MTF.tests, RLE.tests, FMIndex.tests])
  if errors c + failures c == 0 then exitSuccess else exitFailure
